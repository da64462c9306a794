"""GPU parity of the implicit-GEMM tcgen05 convolution kernels (TMA 4-D boxes with out-of-bounds zero
fill as padding) against torch's fp32 convolution of the same bf16 operands."""
import pytest
import torch

from tests.torch_backend import TorchBackend
from tests.test_kernels_gpu import close, rnd

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    return torch.device("cuda", 0)


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(2, 12, 9, 64, 64), (2, 64, 83, 64, 64), (3, 33, 41, 64, 128),
                                            (2, 32, 41, 128, 128), (1, 5, 130, 128, 64), (1, 1, 1, 64, 64)])
def test_umma_conv3x3(dev, B, H, W, Cin, Cout):
    from metaasr_crossaccent_b200.ops import CudaBackend
    bf = torch.bfloat16
    cb, tb = CudaBackend(dev, bf, gemm="umma"), TorchBackend(dev, bf)
    x = rnd((B, H, W, Cin), dev, bf, 1)
    w = rnd((Cout, Cin, 3, 3), dev, torch.float32, 2, 0.05)
    bias = rnd((Cout,), dev, torch.float32, 3)
    wp = torch.empty(Cout, 9 * Cin, device=dev, dtype=bf)
    cb.conv_w_prep(w, wp)
    before = cb.launches
    y1 = torch.full((B, H, W, Cout), 9.0, device=dev, dtype=bf)
    y2 = torch.empty_like(y1)
    cb.conv3x3_fwd(x, wp, bias, y1)
    assert cb.launches == before + 1                      # one implicit-GEMM kernel, no im2col
    tb.conv3x3_fwd(x, wp, bias, y2)
    close(y1, y2, bf, what="umma conv fwd")
    dy = rnd((B, H, W, Cout), dev, bf, 4)
    wpt = torch.empty(Cin, 9 * Cout, device=dev, dtype=bf)
    cb.conv_w_prep_t(w, wpt)
    assert torch.equal(wpt.view(Cin, 9, Cout), wp.view(Cout, 9, Cin).permute(2, 1, 0))
    for mask in (None, x):
        for t in (None, wpt):              # weights read MN-major from wp, or K-major from the transposed layout
            dx1 = torch.full((B, H, W, Cin), 9.0, device=dev, dtype=bf)
            dx2 = torch.empty_like(dx1)
            cb.conv3x3_dgrad(dy, wp, dx1, mask, wpt=t)
            tb.conv3x3_dgrad(dy, wp, dx2, mask)
            close(dx1, dx2, bf, what="umma conv dgrad")
    dwp1 = rnd((Cout, 9 * Cin), dev, torch.float32, 5)
    db1 = torch.zeros(Cout, device=dev)
    dwp2, db2 = dwp1.clone(), db1.clone()
    cb.conv3x3_wgrad(x, dy, dwp1, db1)
    tb.conv3x3_wgrad(x, dy, dwp2, db2)
    close(dwp1, dwp2, torch.float32, 2e-3, "umma conv wgrad")
    close(db1, db2, torch.float32, 1e-4, "conv bias grad")


@pytest.mark.parametrize("B,H,W", [(2, 13, 83), (3, 64, 83), (1, 1, 1), (2, 100, 37)])
def test_umma_conv1(dev, B, H, W):
    """First convolution (Cin = 1) on the tensor cores: patches built in shared memory as a UMMA operand (forward),
    dY streamed by TMA with K = pixels and an all-ones tap row for the bias gradient (wgrad)."""
    from metaasr_crossaccent_b200.ops import CudaBackend
    bf = torch.bfloat16
    cb, tb = CudaBackend(dev, bf, gemm="umma"), TorchBackend(dev, bf)
    x = rnd((B, H, W), dev, torch.float32, 1)
    w = rnd((64, 1, 3, 3), dev, torch.float32, 2, 0.3)
    bias = rnd((64,), dev, torch.float32, 3, 0.1)
    y1 = torch.full((B, H, W, 64), 9.0, device=dev, dtype=bf)
    y2 = torch.empty_like(y1)
    cb.conv1_fwd(x, w, bias, y1)
    tb.conv1_fwd(x, w, bias, y2)
    close(y1, y2, bf, what="umma conv1 fwd")
    dy = rnd((B, H, W, 64), dev, bf, 4) * (y2 > 0)
    dw1, db1 = rnd((64, 1, 3, 3), dev, torch.float32, 5), rnd((64,), dev, torch.float32, 6)
    dw2, db2 = dw1.clone(), db1.clone()
    cb.conv1_wgrad(x, dy, dw1, db1)
    tb.conv1_wgrad(x, dy, dw2, db2)
    close(dw1, dw2, torch.float32, 1e-2, "umma conv1 wgrad")      # x enters the tensor core as bf16
    close(db1, db2, torch.float32, 1e-4, "umma conv1 bias grad")  # exact products (dy * 1.0), fp32 accumulation
