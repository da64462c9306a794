"""GPU parity of the tcgen05 attention kernels (head dim 64, bf16) against the fp32 torch contract."""
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


@pytest.mark.parametrize("B,H,Lq,Lk,causal,use_klens,bwd", [
    (2, 8, 128, 128, False, True, True),      # encoder self-attention at the BASELINE shape
    (3, 8, 33, 33, True, False, True),        # decoder causal self-attention
    (3, 8, 33, 128, False, True, True),       # decoder cross-attention with memory key padding
    (2, 4, 200, 100, False, True, True),      # two query tiles
    (2, 4, 70, 300, False, True, True),       # three key tiles (forward: online soft-max; backward: dQ partials reduced in fp32)
    (1, 2, 150, 150, True, False, True),      # causal over two tiles
    (4, 8, 375, 375, False, True, True),      # encoder self-attention at max_ilen 1500 (fometa-hkust.yaml:45-47): T' = 375
    (4, 8, 33, 375, False, True, True),       # decoder cross-attention over a 375-frame memory
])
def test_umma_attention(dev, B, H, Lq, Lk, causal, use_klens, bwd):
    from metaasr_crossaccent_b200.ops import CudaBackend
    bf = torch.bfloat16
    cb, tb = CudaBackend(dev, bf, gemm="umma"), TorchBackend(dev, bf)
    d = H * 64
    if Lq == Lk:
        qkv = rnd((B * Lq, 3 * d), dev, bf, 1)
        q, k, v = qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:]
    else:
        q = rnd((B * Lq, d), dev, bf, 1)
        kv = rnd((B * Lk, 2 * d), dev, bf, 2)
        k, v = kv[:, :d], kv[:, d:]
    klens = None
    if use_klens:
        klens = torch.tensor([Lk] + [max(1, Lk - 17 * (i + 1)) for i in range(B - 1)], dtype=torch.int64, device=dev)
    o1 = torch.full((B * Lq, d), 5.0, device=dev, dtype=bf)
    o2 = torch.empty_like(o1)
    l1 = torch.empty(B * H * Lq, device=dev)
    l2 = torch.empty_like(l1)
    before = cb.launches
    cb.attn_fwd(q, k, v, o1, l1, B, H, Lq, Lk, klens, causal)
    assert cb.launches == before + 1
    tb.attn_fwd(q, k, v, o2, l2, B, H, Lq, Lk, klens, causal)
    close(o1, o2, bf, what="umma attn out")
    close(l1, l2, bf, what="umma attn lse")
    if not bwd:
        return
    do = rnd((B * Lq, d), dev, bf, 3)
    if Lq == Lk:
        g1 = torch.full((B * Lq, 3 * d), 3.0, device=dev, dtype=bf)
        g2 = g1.clone()
        v1 = (g1[:, :d], g1[:, d:2 * d], g1[:, 2 * d:])
        v2 = (g2[:, :d], g2[:, d:2 * d], g2[:, 2 * d:])
    else:
        v1 = (torch.full((B * Lq, d), 3.0, device=dev, dtype=bf), torch.full((B * Lk, d), 3.0, device=dev, dtype=bf),
              torch.full((B * Lk, d), 3.0, device=dev, dtype=bf))
        v2 = tuple(t.clone() for t in v1)
    dsum = torch.empty(B * H * Lq, device=dev)
    cb.attn_bwd(q, k, v, o2, do, l2, dsum, v1[0], v1[1], v1[2], B, H, Lq, Lk, klens, causal)
    tb.attn_bwd(q, k, v, o2, do, l2, dsum, v2[0], v2[1], v2[2], B, H, Lq, Lk, klens, causal)
    for a, b_, nm in zip(v1, v2, "qkv"):
        close(a, b_, bf, what=f"umma attn d{nm}")


def test_umma_attention_dropout(dev):
    from metaasr_crossaccent_b200.ops import CudaBackend
    bf = torch.bfloat16
    cb = CudaBackend(dev, bf, gemm="umma")
    B, H, L = 2, 4, 128
    d = H * 64
    q = torch.zeros(B * L, d, device=dev, dtype=bf)
    k = torch.zeros(B * L, d, device=dev, dtype=bf)
    v = torch.ones(B * L, d, device=dev, dtype=bf)
    o = torch.empty(B * L, d, device=dev, dtype=bf)
    lse = torch.empty(B * H * L, device=dev)
    cb.attn_fwd(q, k, v, o, lse, B, H, L, L, None, False, 0.25, 99, 3)
    assert abs(float(o.float().mean()) - 1.0) < 0.02          # E[kept / (1-p)] = 1
    o2 = torch.empty_like(o)
    cb.attn_fwd(q, k, v, o2, lse, B, H, L, L, None, False, 0.25, 99, 3)
    assert torch.equal(o, o2)
    # backward replays the same mask: dV_j = sum_i P_ij M_ij dO_i with uniform P -> mean 1 for dO = 1
    do = torch.ones_like(o)
    dq, dk, dv = torch.empty_like(o), torch.empty_like(o), torch.empty_like(o)
    dsum = torch.empty(B * H * L, device=dev)
    cb.attn_bwd(q, k, v, o, do, lse, dsum, dq, dk, dv, B, H, L, L, None, False, 0.25, 99, 3)
    assert abs(float(dv.float().mean()) - 1.0) < 0.02


@pytest.mark.parametrize("B,H,Lq,Lk,causal", [(2, 8, 128, 128, False), (3, 8, 33, 33, True), (3, 8, 33, 128, False)])
def test_umma_attention_dropout_matches_cuda_core_kernels(dev, B, H, Lq, Lk, causal):
    """The tcgen05 and the CUDA-core attention kernels share ONE dropout function (common.cuh attn_drop_*), so with
    the same (seed, site) they drop the same probabilities: forward and backward agree to bf16 rounding."""
    from metaasr_crossaccent_b200.ops import CudaBackend
    bf = torch.bfloat16
    cu, cs = CudaBackend(dev, bf, gemm="umma"), CudaBackend(dev, bf, gemm="simt")
    d = H * 64
    q = rnd((B * Lq, d), dev, bf, 1)
    kv = rnd((B * Lk, 2 * d), dev, bf, 2)
    k, v = kv[:, :d], kv[:, d:]
    klens = None if causal else torch.tensor([Lk] + [max(1, Lk - 9 * (i + 1)) for i in range(B - 1)], dtype=torch.int64, device=dev)
    outs = []
    for be in (cu, cs):
        o = torch.empty(B * Lq, d, device=dev, dtype=bf)
        lse = torch.empty(B * H * Lq, device=dev)
        be.attn_fwd(q, k, v, o, lse, B, H, Lq, Lk, klens, causal, 0.3, 1234, 7)
        do = rnd((B * Lq, d), dev, bf, 3)
        dq, dk, dv = torch.empty_like(q), torch.empty(B * Lk, d, device=dev, dtype=bf), torch.empty(B * Lk, d, device=dev, dtype=bf)
        dsum = torch.empty(B * H * Lq, device=dev)
        be.attn_bwd(q, k, v, o, do, lse, dsum, dq, dk, dv, B, H, Lq, Lk, klens, causal, 0.3, 1234, 7)
        outs.append((o, lse, dq, dk, dv))
    for a, b_, nm in zip(outs[0], outs[1], ("out", "lse", "dq", "dk", "dv")):
        close(a, b_, bf, what=f"dropout parity {nm}")


@pytest.mark.parametrize("B,H,Lq,Lk,causal,use_klens", [
    (3, 8, 33, 33, True, False),         # decoder causal self-attention (benchmark shape)
    (3, 8, 33, 128, False, True),        # decoder cross-attention, two key tiles of the warp-MMA kernel
    (2, 8, 64, 64, True, False),         # the longest query sequence the warp-MMA kernels take
    (2, 4, 50, 200, False, True),        # ragged: 4 key tiles, the last one partial, rows 48-49 in the fourth warp
    (5, 2, 1, 17, False, False),         # one decode row
    (2, 8, 16, 375, False, True),        # a single active warp over a 375-frame memory
])
@pytest.mark.parametrize("family", ["warp_mma", "tcgen05"])
def test_short_query_attention_both_families(dev, B, H, Lq, Lk, causal, use_klens, family):
    """Query sequences of <= 64 rows run on the warp-MMA kernels (attn_small.cu) by default; masr_attn_set_small_lq(0)
    sends them to the tcgen05 kernels.  Both families against the fp32 torch contract, forward and backward, with the D rows
    computed in the kernel and supplied by the caller (as the out-projection dgrad does), and with dropout replay."""
    from metaasr_crossaccent_b200.ops import CudaBackend
    bf = torch.bfloat16
    cb, tb = CudaBackend(dev, bf, gemm="umma"), TorchBackend(dev, bf)
    cb.set_attn_small_lq(64 if family == "warp_mma" else 0)
    try:
        d = H * 64
        q = rnd((B * Lq, d), dev, bf, 1)
        kv = rnd((B * Lk, 2 * d), dev, bf, 2)
        k, v = kv[:, :d], kv[:, d:]
        klens = None
        if use_klens:
            klens = torch.tensor([Lk] + [max(1, Lk - 23 * (i + 1)) for i in range(B - 1)], dtype=torch.int64, device=dev)
        o1, o2 = torch.full((B * Lq, d), 5.0, device=dev, dtype=bf), torch.empty(B * Lq, d, device=dev, dtype=bf)
        l1, l2 = torch.empty(B * H * Lq, device=dev), torch.empty(B * H * Lq, device=dev)
        cb.attn_fwd(q, k, v, o1, l1, B, H, Lq, Lk, klens, causal)
        tb.attn_fwd(q, k, v, o2, l2, B, H, Lq, Lk, klens, causal)
        close(o1, o2, bf, what=f"{family} out")
        close(l1, l2, bf, what=f"{family} lse")
        do = rnd((B * Lq, d), dev, bf, 3)
        ref = (torch.empty_like(q), torch.empty(B * Lk, d, device=dev, dtype=bf), torch.empty(B * Lk, d, device=dev, dtype=bf))
        dsum = torch.empty(B * H * Lq, device=dev)
        tb.attn_bwd(q, k, v, o2, do, l2, dsum, ref[0], ref[1], ref[2], B, H, Lq, Lk, klens, causal)
        D = (do.float() * o2.float()).view(B, Lq, H, 64).sum(-1).permute(0, 2, 1).contiguous().view(-1)
        for ready in (False, True):
            got = (torch.full_like(q, 3.0), torch.full((B * Lk, d), 3.0, device=dev, dtype=bf),
                   torch.full((B * Lk, d), 3.0, device=dev, dtype=bf))
            ds = D.clone() if ready else torch.empty(B * H * Lq, device=dev)
            cb.attn_bwd(q, k, v, o2, do, l2, ds, got[0], got[1], got[2], B, H, Lq, Lk, klens, causal, dsum_ready=ready)
            for a, b_, nm in zip(got, ref, "qkv"):
                close(a, b_, bf, what=f"{family} d{nm} (dsum_ready={ready})")
        # dropout: the same (seed, site) drops the same probabilities in both families and in the CUDA-core kernels
        cs = CudaBackend(dev, bf, gemm="simt")
        outs = []
        for be in (cb, cs):
            o = torch.empty(B * Lq, d, device=dev, dtype=bf)
            lse = torch.empty(B * H * Lq, device=dev)
            be.attn_fwd(q, k, v, o, lse, B, H, Lq, Lk, klens, causal, 0.3, 1234, 7)
            g = (torch.empty_like(q), torch.empty(B * Lk, d, device=dev, dtype=bf), torch.empty(B * Lk, d, device=dev, dtype=bf))
            be.attn_bwd(q, k, v, o, do, lse, torch.empty(B * H * Lq, device=dev), g[0], g[1], g[2], B, H, Lq, Lk, klens, causal,
                        0.3, 1234, 7)
            outs.append((o, lse) + g)
        for a, b_, nm in zip(outs[0], outs[1], ("out", "lse", "dq", "dk", "dv")):
            close(a, b_, bf, what=f"{family} dropout parity {nm}")
    finally:
        cb.set_attn_small_lq(64)


def test_short_query_attention_decode_cache(dev):
    """kv_rows > Lk: one decode row per utterance against a key/value cache of fixed capacity, read in place."""
    from metaasr_crossaccent_b200.ops import CudaBackend
    bf = torch.bfloat16
    cb, tb = CudaBackend(dev, bf, gemm="umma"), TorchBackend(dev, bf)
    B, H, cap, Lk = 4, 8, 40, 23
    d = H * 64
    q = rnd((B, d), dev, bf, 1)
    kc, vc = rnd((B * cap, d), dev, bf, 2), rnd((B * cap, d), dev, bf, 3)
    o1, o2 = torch.empty(B, d, device=dev, dtype=bf), torch.empty(B, d, device=dev, dtype=bf)
    l1, l2 = torch.empty(B * H, device=dev), torch.empty(B * H, device=dev)
    cb.attn_fwd(q, kc, vc, o1, l1, B, H, 1, Lk, None, False, kv_rows=cap)
    kd = kc.view(B, cap, d)[:, :Lk].reshape(B * Lk, d).contiguous()
    vd = vc.view(B, cap, d)[:, :Lk].reshape(B * Lk, d).contiguous()
    tb.attn_fwd(q, kd, vd, o2, l2, B, H, 1, Lk, None, False)
    close(o1, o2, bf, what="cached out")
    close(l1, l2, bf, what="cached lse")
