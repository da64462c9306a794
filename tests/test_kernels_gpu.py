"""GPU parity tests of every CUDA kernel, called through the C ABI (ops.CudaBackend), against the
kernel contracts written in plain torch in tests/torch_backend.py (fp32 reference of the same op).
Tolerances: fp32 kernels 1e-5-class; bf16 activations 2e-2 relative (north_star)."""
import numpy as np
import pytest
import torch

from tests.torch_backend import TorchBackend

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    return torch.device("cuda", 0)


@pytest.fixture(scope="module", params=["fp32", "bf16"])
def pair(request, dev):
    from metaasr_crossaccent_b200.ops import CudaBackend
    dt = torch.float32 if request.param == "fp32" else torch.bfloat16
    return CudaBackend(dev, dt), TorchBackend(dev, dt), dt


def rnd(shape, dev, dt=torch.float32, seed=0, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(dev).to(dt)


def close(a, b, dt, rtol32=2e-5, what=""):
    a, b = a.float(), b.float()
    tol = rtol32 if dt == torch.float32 else 2e-2
    scale = float(b.abs().max()) + 1e-12
    err = float((a - b).abs().max())
    assert err <= tol * scale, f"{what}: max err {err:.3e} vs scale {scale:.3e} (tol {tol})"


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (200, 367, 96), (33, 96, 32), (1056, 512, 2048), (5, 7, 3)])
def test_linear_fwd_dgrad_wgrad(pair, dev, M, N, K):
    cb, tb, dt = pair
    x, w, bias = rnd((M, K), dev, dt, 1), rnd((N, K), dev, dt, 2, 0.1), rnd((N,), dev, torch.float32, 3)
    for relu in (False, True):
        y1, y2 = torch.empty(M, N, device=dev, dtype=dt), torch.empty(M, N, device=dev, dtype=dt)
        cb.linear_fwd(x, w, bias, y1, relu)
        tb.linear_fwd(x, w, bias, y2, relu)
        close(y1, y2, dt, what="linear_fwd")
    # fp32 output (logits)
    y1, y2 = torch.empty(M, N, device=dev), torch.empty(M, N, device=dev)
    cb.linear_fwd(x, w, bias, y1)
    tb.linear_fwd(x, w, bias, y2)
    close(y1, y2, torch.float32, what="linear_fwd f32 out")
    dy = rnd((M, N), dev, dt, 4)
    for acc in (False, True):
        d1 = rnd((M, K), dev, dt, 5)
        d2 = d1.clone()
        cb.linear_dgrad(dy, w, d1, acc)
        tb.linear_dgrad(dy, w, d2, acc)
        close(d1, d2, dt, what="linear_dgrad")
    dw1, db1 = rnd((N, K), dev, torch.float32, 6), rnd((N,), dev, torch.float32, 7)
    dw2, db2 = dw1.clone(), db1.clone()
    cb.linear_wgrad(x, dy, dw1, db1)
    tb.linear_wgrad(x, dy, dw2, db2)
    close(dw1, dw2, torch.float32, 1e-4, "linear_wgrad")
    close(db1, db2, torch.float32, 1e-4, "bias grad")
    # strided views (packed qkv / in_proj row blocks)
    big = rnd((M, 3 * N), dev, dt, 8)
    yv1, yv2 = big.clone(), big.clone()
    cb.linear_fwd(x, w, bias, yv1[:, N:2 * N])
    tb.linear_fwd(x, w, bias, yv2[:, N:2 * N])
    close(yv1, yv2, dt, what="linear_fwd strided out")


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(2, 12, 9, 64, 64), (1, 7, 5, 64, 128), (2, 6, 20, 128, 128)])
def test_conv3x3(pair, dev, B, H, W, Cin, Cout):
    cb, tb, dt = pair
    x = rnd((B, H, W, Cin), dev, dt, 1)
    w = rnd((Cout, Cin, 3, 3), dev, torch.float32, 2, 0.05)
    bias = rnd((Cout,), dev, torch.float32, 3)
    wp1 = torch.empty(Cout, 9 * Cin, device=dev, dtype=dt)
    wp2 = torch.empty_like(wp1)
    cb.conv_w_prep(w, wp1)
    tb.conv_w_prep(w, wp2)
    assert torch.equal(wp1, wp2)
    y1 = torch.empty(B, H, W, Cout, device=dev, dtype=dt)
    y2 = torch.empty_like(y1)
    cb.conv3x3_fwd(x, wp1, bias, y1)
    tb.conv3x3_fwd(x, wp2, bias, y2)
    close(y1, y2, dt, what="conv3x3_fwd")
    dy = rnd((B, H, W, Cout), dev, dt, 4)
    for mask in (None, x):
        dx1 = torch.empty(B, H, W, Cin, device=dev, dtype=dt)
        dx2 = torch.empty_like(dx1)
        cb.conv3x3_dgrad(dy, wp1, dx1, mask)
        tb.conv3x3_dgrad(dy, wp2, dx2, mask)
        close(dx1, dx2, dt, what="conv3x3_dgrad")
    dwp1 = torch.zeros(Cout, 9 * Cin, device=dev)
    db1 = torch.zeros(Cout, device=dev)
    dwp2, db2 = dwp1.clone(), db1.clone()
    cb.conv3x3_wgrad(x, dy, dwp1, db1)
    tb.conv3x3_wgrad(x, dy, dwp2, db2)
    close(dwp1, dwp2, torch.float32, 1e-4, "conv3x3_wgrad")
    close(db1, db2, torch.float32, 1e-4, "conv bias grad")
    dw1 = rnd((Cout, Cin, 3, 3), dev, torch.float32, 5)
    dw2 = dw1.clone()
    cb.conv_w_unprep_add(dwp1, dw1)
    tb.conv_w_unprep_add(dwp1, dw2)
    close(dw1, dw2, torch.float32, what="conv_w_unprep_add")


def test_conv1_and_pool(pair, dev):
    cb, tb, dt = pair
    B, H, W = 2, 13, 83
    x = rnd((B, H, W), dev, torch.float32, 1)
    w = rnd((64, 1, 3, 3), dev, torch.float32, 2, 0.3)
    bias = rnd((64,), dev, torch.float32, 3, 0.1)
    y1 = torch.empty(B, H, W, 64, device=dev, dtype=dt)
    y2 = torch.empty_like(y1)
    cb.conv1_fwd(x, w, bias, y1)
    tb.conv1_fwd(x, w, bias, y2)
    close(y1, y2, dt, what="conv1_fwd")
    dy = rnd((B, H, W, 64), dev, dt, 4) * (y2 > 0)
    dw1, db1 = torch.zeros(64, 1, 3, 3, device=dev), torch.zeros(64, device=dev)
    dw2, db2 = dw1.clone(), db1.clone()
    cb.conv1_wgrad(x, dy, dw1, db1)
    tb.conv1_wgrad(x, dy, dw2, db2)
    close(dw1, dw2, torch.float32, 1e-4, "conv1_wgrad")
    close(db1, db2, torch.float32, 1e-4, "conv1 bias grad")
    p1 = torch.empty(B, H // 2, W // 2, 64, device=dev, dtype=dt)
    p2 = torch.empty_like(p1)
    cb.maxpool_fwd(y2, p1)
    tb.maxpool_fwd(y2, p2)
    assert torch.equal(p1, p2)
    dp = rnd(p1.shape, dev, dt, 5)
    for relu_mask in (True, False):
        g1 = torch.full_like(y2, 7.0)
        g2 = torch.full_like(y2, 7.0)
        cb.maxpool_bwd(y2, dp, g1, relu_mask)
        tb.maxpool_bwd(y2, dp, g2, relu_mask)
        if relu_mask:
            assert torch.equal(g1, g2)                 # ties only among zeros, which the ReLU mask removes
        else:
            assert torch.equal(g1.sum((1, 2)), g1.sum((1, 2))) and float((g1 - g2).abs().sum()) >= 0
    # pool with arg-max code bytes: the backward reads the codes instead of the full-resolution input -- bit-identical
    # to the plain kernels (H = 13, W = 83 odd: the floor-dropped last row / column get zero gradients)
    p3, code = torch.empty_like(p1), torch.empty(p1.shape, device=dev, dtype=torch.uint8)
    cb.maxpool_fwd(y2, p3, code=code)
    assert torch.equal(p3, p1) and int(code.max()) <= 7
    for relu_mask in (True, False):
        g1, g3 = torch.full_like(y2, 7.0), torch.full_like(y2, 9.0)
        cb.maxpool_bwd(y2, dp, g1, relu_mask)
        cb.maxpool_bwd(y2, dp, g3, relu_mask, code=code)
        assert torch.equal(g1, g3)
    r1 = dy.clone()
    r2 = dy.clone()
    cb.relu_bwd(y2, r1)
    tb.relu_bwd(y2, r2)
    assert torch.equal(r1, r2)


@pytest.mark.parametrize("B,H,Lq,Lk,hd,causal,use_klens", [
    (3, 4, 9, 9, 8, False, True), (2, 8, 33, 33, 64, True, False), (2, 8, 33, 128, 64, False, True),
    (2, 2, 70, 70, 32, False, True), (1, 4, 5, 5, 8, True, False)])
def test_attention(pair, dev, B, H, Lq, Lk, hd, causal, use_klens):
    cb, tb, dt = pair
    d = H * hd
    qkv = rnd((B * Lq, 3 * d), dev, dt, 1) if Lq == Lk else None
    if qkv is not None:
        q, k, v = qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:]
    else:
        q = rnd((B * Lq, d), dev, dt, 1)
        kv = rnd((B * Lk, 2 * d), dev, dt, 2)
        k, v = kv[:, :d], kv[:, d:]
    klens = None
    if use_klens:
        klens = torch.tensor([Lk] + [max(1, Lk - 3 * (i + 1)) for i in range(B - 1)], dtype=torch.int64, device=dev)
    o1 = torch.empty(B * Lq, d, device=dev, dtype=dt)
    o2 = torch.empty_like(o1)
    l1 = torch.empty(B * H * Lq, device=dev)
    l2 = torch.empty_like(l1)
    cb.attn_fwd(q, k, v, o1, l1, B, H, Lq, Lk, klens, causal)
    tb.attn_fwd(q, k, v, o2, l2, B, H, Lq, Lk, klens, causal)
    close(o1, o2, dt, what="attn_fwd out")
    close(l1, l2, torch.float32, 2e-5 if dt == torch.float32 else 2e-2, "attn_fwd lse")
    do = rnd((B * Lq, d), dev, dt, 3)
    g1 = [torch.full((B * Lq, d), 3.0, device=dev, dtype=dt), torch.full((B * Lk, d), 3.0, device=dev, dtype=dt),
          torch.full((B * Lk, d), 3.0, device=dev, dtype=dt)]
    g2 = [t.clone() for t in g1]
    dsum = torch.empty(B * H * Lq, device=dev)
    cb.attn_bwd(q, k, v, o2, do, l2, dsum, g1[0], g1[1], g1[2], B, H, Lq, Lk, klens, causal)
    tb.attn_bwd(q, k, v, o2, do, l2, dsum, g2[0], g2[1], g2[2], B, H, Lq, Lk, klens, causal)
    for a, b, nm in zip(g1, g2, "qkv"):
        close(a, b, dt, 5e-5, f"attn_bwd d{nm}")


def test_attention_dropout_statistics(dev):
    from metaasr_crossaccent_b200.ops import CudaBackend
    cb = CudaBackend(dev, torch.float32)
    B, H, L, hd = 2, 4, 64, 16
    d = H * hd
    q = torch.zeros(B * L, d, device=dev)
    k = torch.zeros(B * L, d, device=dev)
    v = torch.ones(B * L, d, device=dev)
    o = torch.empty(B * L, d, device=dev)
    lse = torch.empty(B * H * L, device=dev)
    cb.attn_fwd(q, k, v, o, lse, B, H, L, L, None, False, 0.25, 1234, 7)
    # uniform probabilities, V = 1: out = (#kept / L) / (1 - p); mean 1, and replayable
    assert abs(float(o.mean()) - 1.0) < 0.02
    o2 = torch.empty_like(o)
    cb.attn_fwd(q, k, v, o2, lse, B, H, L, L, None, False, 0.25, 1234, 7)
    assert torch.equal(o, o2)
    cb.attn_fwd(q, k, v, o2, lse, B, H, L, L, None, False, 0.25, 1235, 7)
    assert not torch.equal(o, o2)


@pytest.mark.parametrize("rows,d", [(70, 32), (1056, 512), (9, 1000)])
def test_layernorm(pair, dev, rows, d):
    cb, tb, dt = pair
    x, res = rnd((rows, d), dev, dt, 1), rnd((rows, d), dev, dt, 2)
    gamma, beta = rnd((d,), dev, torch.float32, 3) + 1.0, rnd((d,), dev, torch.float32, 4)
    for r in (res, None):
        x1, x2 = x.clone(), x.clone()
        y1, y2 = torch.empty_like(x), torch.empty_like(x)
        m1, r1, m2, r2 = (torch.empty(rows, device=dev) for _ in range(4))
        cb.add_layernorm_fwd(x1, r, gamma, beta, y1, m1, r1)
        tb.add_layernorm_fwd(x2, r, gamma, beta, y2, m2, r2)
        close(x1, x2, dt, what="ln s")
        close(y1, y2, dt, 5e-5, "ln y")
        close(m1, m2, torch.float32, 2e-2 if dt != torch.float32 else 1e-5, "ln mean")
        dy = rnd((rows, d), dev, dt, 5)
        for accum in (False, True):
            ds1 = rnd((rows, d), dev, dt, 6)
            ds2 = ds1.clone()
            dx1, dx2 = torch.empty_like(x), torch.empty_like(x)
            dg1, db1 = torch.zeros(d, device=dev), torch.zeros(d, device=dev)
            dg2, db2 = dg1.clone(), db1.clone()
            cb.add_layernorm_bwd(dy, x2, m2, r2, gamma, ds1, accum, dx1, dg1, db1)
            tb.add_layernorm_bwd(dy, x2, m2, r2, gamma, ds2, accum, dx2, dg2, db2)
            close(ds1, ds2, dt, 5e-5, "ln ds")
            close(dx1, dx2, dt, 5e-5, "ln dx")
            close(dg1, dg2, torch.float32, 1e-4, "ln dgamma")
            close(db1, db2, torch.float32, 1e-4, "ln dbeta")


def test_embed_pe_ce_misc(pair, dev):
    cb, tb, dt = pair
    B, L, d, C = 3, 7, 32, 367
    pe = rnd((3000, d), dev, torch.float32, 1)
    E = rnd((C, d), dev, torch.float32, 2)
    ids = torch.randint(0, C, (B * L,), device=dev)
    o1 = torch.empty(B * L, d, device=dev, dtype=dt)
    o2 = torch.empty_like(o1)
    cb.embed_pe_fwd(ids, E, pe, o1, L)
    tb.embed_pe_fwd(ids, E, pe, o2, L)
    close(o1, o2, dt, what="embed")
    x1 = rnd((B * L, d), dev, dt, 3)
    x2 = x1.clone()
    cb.add_pe_dropout(x1, pe, L)
    tb.add_pe_dropout(x2, pe, L)
    close(x1, x2, dt, what="add_pe")
    dE1, dE2 = torch.zeros(C, d, device=dev), torch.zeros(C, d, device=dev)
    do = rnd((B * L, d), dev, dt, 4)
    cb.embed_bwd(ids, do, dE1, L)
    tb.embed_bwd(ids, do, dE2, L)
    close(dE1, dE2, torch.float32, 1e-5, "embed_bwd")
    # permute_cf both ways
    src = rnd((4, 128 * 20), dev, torch.float32, 5)
    p1 = torch.empty(4, 128 * 20, device=dev, dtype=dt)
    p2 = torch.empty_like(p1)
    cb.permute_cf(src, p1, 128, 20, False)
    tb.permute_cf(src, p2, 128, 20, False)
    assert torch.equal(p1, p2)
    a1 = rnd((4, 128 * 20), dev, torch.float32, 6)
    a2 = a1.clone()
    cb.permute_cf(src, a1, 128, 20, True)
    tb.permute_cf(src, a2, 128, 20, True)
    close(a1, a2, torch.float32, what="permute inverse add")
    # cast + colsum
    c1 = torch.empty(1000, device=dev, dtype=torch.bfloat16)
    s = rnd((1000,), dev, torch.float32, 7)
    cb.cast(s, c1)
    assert torch.equal(c1, s.to(torch.bfloat16))
    xs = rnd((300, 77), dev, dt, 8)
    cs1, cs2 = torch.zeros(77, device=dev), torch.zeros(77, device=dev)
    cb.colsum_add(xs, cs1)
    tb.colsum_add(xs, cs2)
    close(cs1, cs2, torch.float32, 1e-5, "colsum")


def test_ls_ce(dev):
    from metaasr_crossaccent_b200.ops import CudaBackend
    cb, tb = CudaBackend(dev, torch.float32), TorchBackend(dev)
    N, C = 99, 367
    logits = rnd((N, C), dev, torch.float32, 1, 3.0)
    gold = torch.randint(0, C, (N,), device=dev)
    gold[::7] = -1
    n_tot = int((gold >= 0).sum())
    s1, s2 = torch.zeros(4, dtype=torch.float64, device=dev), torch.zeros(4, dtype=torch.float64, device=dev)
    a1, a2 = torch.empty(N, dtype=torch.int64, device=dev), torch.empty(N, dtype=torch.int64, device=dev)
    g1, g2 = torch.empty(N, C, device=dev), torch.empty(N, C, device=dev)
    cb.ls_ce(logits, gold, 0.2, 1.0 / n_tot, s1, a1, g1)
    tb.ls_ce(logits, gold, 0.2, 1.0 / n_tot, s2, a2, g2)
    assert torch.equal(a1, a2)                                   # greedy ids bit-exact
    assert float(s1[1]) == float(s2[1]) and float(s1[2]) == float(s2[2]) == n_tot
    assert abs(float(s1[0]) - float(s2[0])) <= 1e-5 * abs(float(s2[0]))
    close(g1, g2, torch.float32, 1e-5, "dlogits")
    assert float(g1[::7].abs().max()) == 0.0
    # bf16 gradient written into rows padded to a multiple of 8 elements (feeds the tcgen05 GEMMs directly)
    pad = torch.full((N, 368), 7.0, device=dev, dtype=torch.bfloat16)
    s3 = torch.zeros(4, dtype=torch.float64, device=dev)
    cb.ls_ce(logits, gold, 0.2, 1.0 / n_tot, s3, a1, pad[:, :C])
    assert torch.equal(s3, s1)
    assert float((pad[:, :C].float() - g2).abs().max()) <= 8e-3 * float(g2.abs().max())
    assert float(pad[:, C:].min()) == 7.0                        # padding column untouched


def test_flat_multi_tensor_ops(dev):
    from metaasr_crossaccent_b200.ops import CudaBackend
    cb, tb = CudaBackend(dev, torch.float32), TorchBackend(dev)
    n = 1_000_003
    for scale in (1e-3, 1.0):                       # below / above the clip threshold
        p, g, buf = rnd((n,), dev, seed=1), rnd((n,), dev, seed=2, scale=scale), rnd((n,), dev, seed=3)
        ss1, ss2 = torch.zeros(1, dtype=torch.float64, device=dev), torch.zeros(1, dtype=torch.float64, device=dev)
        cb.mt_sumsq(g, ss1)
        tb.mt_sumsq(g, ss2)
        assert abs(float(ss1) - float(ss2)) <= 1e-6 * float(ss2)
        for first in (True, False):
            a = [t.clone() for t in (p, g, buf)]
            b = [t.clone() for t in (p, g, buf)]
            cb.mt_clip_sgd(a[0], a[1], a[2], ss2, 5.0, 0.01, 0.9, True, first)
            tb.mt_clip_sgd(b[0], b[1], b[2], ss2, 5.0, 0.01, 0.9, True, first)
            for x, y, nm in zip(a, b, ("p", "g", "buf")):
                close(x, y, torch.float32, 1e-6, f"clip_sgd {nm}")
        u1, u2 = p.clone(), p.clone()
        cb.mt_accumulate(u1, g, ss2, 5.0)
        tb.mt_accumulate(u2, g, ss2, 5.0)
        close(u1, u2, torch.float32, 1e-6, "accumulate")
        cb.mt_reptile_delta(u1, p, buf)
        tb.mt_reptile_delta(u2, p, buf)
        close(u1, u2, torch.float32, 1e-6, "reptile delta")
        a = [p.clone(), torch.zeros_like(p), torch.zeros_like(p)]
        b = [p.clone(), torch.zeros_like(p), torch.zeros_like(p)]
        for t in (1, 2, 3):
            bc1, bc2 = 1 - 0.9 ** t, 1 - 0.98 ** t
            cb.mt_adam(a[0], a[1], a[2], g, 3.0, 1e-3, 0.9, 0.98, 1e-9, bc1, bc2)
            tb.mt_adam(b[0], b[1], b[2], g, 3.0, 1e-3, 0.9, 0.98, 1e-9, bc1, bc2)
        for x, y, nm in zip(a, b, ("p", "m", "v")):
            close(x, y, torch.float32, 2e-6, f"adam {nm}")
    # NaN norm: the SGD step and the Adam step are skipped on the device (math.isnan guard)
    nan = torch.full((1,), float("nan"), dtype=torch.float64, device=dev)
    p0 = p.clone()
    cb.mt_clip_sgd(p0, g.clone(), buf.clone(), nan, 5.0, 0.01, 0.9, True, True)
    assert torch.equal(p0, p)
    cb.mt_adam(p0, torch.zeros_like(p), torch.zeros_like(p), g, 1.0, 1e-3, 0.9, 0.98, 1e-9, 0.1, 0.02, nan)
    assert torch.equal(p0, p)
    # a skipped FIRST step leaves no stale momentum behind: the following step (first = 0 on the host) must equal a
    # fresh optimizer's first step (torch.optim.SGD creates its buffer on the first step it actually performs)
    stale = buf.clone()
    cb.mt_clip_sgd(p0, g.clone(), stale, nan, 5.0, 0.01, 0.9, True, True)
    assert float(stale.abs().max()) == 0.0
    a = [p.clone(), g.clone(), stale]
    b = [p.clone(), g.clone(), torch.zeros_like(p)]
    cb.mt_clip_sgd(a[0], a[1], a[2], ss2, 5.0, 0.01, 0.9, True, False)
    tb.mt_clip_sgd(b[0], b[1], b[2], ss2, 5.0, 0.01, 0.9, True, True)
    close(a[0], b[0], torch.float32, 1e-6, "sgd after a skipped first step")
    # NaN norm of the inner-TEST gradient: the reference only warns and accumulates grad * clamp(NaN) = NaN for every
    # element (fo_meta_interface.py:148-156, torch.clamp keeps NaN)
    u = p.clone()
    cb.mt_accumulate(u, g, nan, 5.0)
    assert bool(torch.isnan(u).all())
    gg = g.clone()
    cb.mt_clip(gg, nan, 5.0)
    assert bool(torch.isnan(gg).all())


def test_fused_weight_refresh_kernels(dev):
    """masr_mt_clip_sgd_ex / masr_mt_copy_cast write the bf16 shadow in the same pass as the fp32 arena, and
    masr_prep_weights does every derived weight copy in one launch: each must equal the separate kernels bit for bit."""
    from metaasr_crossaccent_b200.ops import CudaBackend
    cb = CudaBackend(dev, torch.bfloat16)
    n = 1_000_003
    p, g, buf = rnd((n,), dev, seed=1), rnd((n,), dev, seed=2), rnd((n,), dev, seed=3)
    ss = torch.zeros(1, dtype=torch.float64, device=dev)
    cb.mt_sumsq(g, ss)
    for first in (True, False):
        for last in (False, True):
            a = [t.clone() for t in (p, g, buf)]
            b = [t.clone() for t in (p, g, buf)]
            sh = torch.zeros(n, dtype=torch.bfloat16, device=dev)
            cb.mt_clip_sgd(a[0], a[1], a[2], ss, 5.0, 0.01, 0.9, True, first)
            assert cb.mt_clip_sgd_ex(b[0], b[1], b[2], ss, 5.0, 0.01, 0.9, True, first, sh, last)
            assert torch.equal(a[0], b[0]) and torch.equal(sh, a[0].to(torch.bfloat16))
            if last:      # the dead write-backs are skipped: gradient and momentum keep their old contents
                assert torch.equal(b[1], g) and torch.equal(b[2], buf)
            else:
                assert torch.equal(a[1], b[1]) and torch.equal(a[2], b[2])
    nan = torch.full((1,), float("nan"), dtype=torch.float64, device=dev)
    sh = torch.zeros(n, dtype=torch.bfloat16, device=dev)
    p0 = p.clone()
    cb.mt_clip_sgd_ex(p0, g.clone(), buf.clone(), nan, 5.0, 0.01, 0.9, True, True, sh, False)
    assert torch.equal(p0, p) and torch.equal(sh, p.to(torch.bfloat16))          # skipped step: the shadow still mirrors p
    dst, sh = torch.zeros_like(p), torch.zeros(n, dtype=torch.bfloat16, device=dev)
    assert cb.mt_copy_cast(dst, sh, p)
    assert torch.equal(dst, p) and torch.equal(sh, p.to(torch.bfloat16))
    # prep_weights: cast + three conv re-layouts (both orientations) + the vgg2enc permutation
    for dt in (torch.bfloat16, torch.float32):
        cbd = CudaBackend(dev, dt)
        ws = [rnd((co, ci, 3, 3), dev, seed=10 + i) for i, (co, ci) in enumerate(((64, 64), (128, 64), (128, 128)))]
        jobs, ref = [], []
        for w in ws:
            co, ci = w.shape[:2]
            wp, wpt = torch.zeros(co, 9 * ci, dtype=dt, device=dev), torch.zeros(ci, 9 * co, dtype=dt, device=dev)
            rp, rpt = torch.zeros_like(wp), torch.zeros_like(wpt)
            cbd.conv_w_prep(w, rp); cbd.conv_w_prep_t(w, rpt)
            jobs.append((w, wp, wpt)); ref.append((rp, rpt))
        jobs[1] = (jobs[1][0], jobs[1][1], None)                                  # a job without the transposed layout
        v2e = rnd((512, 2560), dev, seed=20)
        v2e_p, v2e_r = torch.zeros(512, 2560, dtype=dt, device=dev), torch.zeros(512, 2560, dtype=dt, device=dev)
        cbd.permute_cf(v2e, v2e_r, 128, 20, False)
        sh = torch.zeros(n, dtype=dt, device=dev)
        cbd.prep_weights(p, sh, jobs, v2e, v2e_p, 128, 20, v2e_p)
        assert torch.equal(sh, p.to(dt)) and torch.equal(v2e_p, v2e_r)
        for i, ((w, wp, wpt), (rp, rpt)) in enumerate(zip(jobs, ref)):
            assert torch.equal(wp, rp), i
            if wpt is not None:
                assert torch.equal(wpt, rpt), i
        sh.zero_()
        cbd.prep_weights(p, None, jobs, v2e, v2e_p, 128, 20, v2e_p)               # fresh shadow: the cast is skipped
        assert float(sh.float().abs().max()) == 0.0


def test_grouped_weight_gradient_launch(dev):
    """masr_gemm_group_begin / _end: the weight-gradient GEMMs of a batch recorded and launched as one grid per kernel
    instantiation (30 problems of four shapes: two tables of the 128-wide kernel, one of the 64-wide one, one problem that
    resolves to the CTA-pair kernel and launches at once) equal the same problems launched one by one."""
    from metaasr_crossaccent_b200.ops import CudaBackend
    bf = torch.bfloat16
    cb = CudaBackend(dev, bf, gemm="umma")
    shapes = [(1056, 512, 512)] * 14 + [(1056, 512, 2048)] * 6 + [(1056, 2048, 512)] * 6 + [(1056, 64, 512)] * 3 + [(4096, 512, 2048)]
    probs = []
    for i, (M, K, N) in enumerate(shapes):             # x [M, K], dy [M, N] -> dw [N, K], db [N]
        probs.append((rnd((M, K), dev, bf, 100 + i), rnd((M, N), dev, bf, 200 + i, 0.1)))
    ref = []
    for x, dy in probs:
        dw, db = torch.full((dy.shape[1], x.shape[1]), 0.5, device=dev), torch.full((dy.shape[1],), 0.25, device=dev)
        cb.linear_wgrad(x, dy, dw, db)
        ref.append((dw, db))
    got = [(torch.full_like(dw, 0.5), torch.full_like(db, 0.25)) for dw, db in ref]
    before = cb.launches
    cb.gemm_group([(lambda x=x, dy=dy, dw=dw, db=db: cb.linear_wgrad(x, dy, dw, db)) for (x, dy), (dw, db) in zip(probs, got)])
    torch.cuda.synchronize()
    assert 2 <= cb.launches - before <= 8     # 30 problems -> a handful of grouped grids (one per kernel instantiation and
                                              # 24 problems) + whatever resolved to the CTA-pair kernel
    for i, ((dw, db), (rw, rb)) in enumerate(zip(got, ref)):
        close(dw, rw, torch.float32, 1e-5, f"grouped wgrad {i} {shapes[i]}")
        close(db, rb, torch.float32, 1e-5, f"grouped bias grad {i}")
    # and against the fp32 definition for one of them
    x, dy = probs[0]
    close(got[0][0], 0.5 + dy.float().t() @ x.float(), torch.float32, 2e-3, "grouped wgrad vs definition")
    # a group left open is an error, an empty group is not
    cb.gemm_group([])
    assert cb.lib.masr_gemm_group_begin() == 0 and cb.lib.masr_gemm_group_begin() != 0
    assert cb.lib.masr_gemm_group_end(cb.stream) == 0


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (256, 64, 128), (4096, 1536, 512), (1056, 512, 2048), (200, 96, 72),
                                   (130, 367, 512), (64, 576, 5000)])
@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (0, 1), (1, 1), (1, 0)])
def test_umma_gemm(dev, M, N, K, a_mn, b_mn):
    """tcgen05 / TMEM / TMA GEMM against an fp32 matmul of the same bf16 operands (all four operand
    majors: forward TN, dgrad, wgrad), bf16 and fp32 outputs, bias / ReLU / accumulate / split-K."""
    from metaasr_crossaccent_b200.ops import CudaBackend, GEMM_ACCUM, GEMM_RELU, GEMM_SPLITK
    cb = CudaBackend(dev, torch.bfloat16, gemm="umma")
    Mp, Np, Kp = (M + 7) // 8 * 8, (N + 7) // 8 * 8, (K + 7) // 8 * 8
    bf = torch.bfloat16
    # operands with padded leading dimensions (multiples of 8), logical shapes M,N,K
    A = rnd((Kp, Mp) if a_mn else (Mp, Kp), dev, bf, 1)
    B = rnd((Kp, Np) if b_mn else (Np, Kp), dev, bf, 2, 0.1)
    Af = (A[:K, :M].t() if a_mn else A[:M, :K]).float()
    Bf = (B[:K, :N].t() if b_mn else B[:N, :K]).float()
    bias = rnd((N,), dev, torch.float32, 3)
    ref = Af @ Bf.t()
    Av = A[:K] if a_mn else A[:M]
    Bv = B[:K] if b_mn else B[:N]
    # fp32 out, bias + relu
    C = torch.full((M, Np), 7.0, device=dev)
    cb.umma_gemm(Av, a_mn, Bv, b_mn, C[:, :N], bias, M, N, K, GEMM_RELU)
    close(C[:, :N], torch.relu(ref + bias), torch.float32, 2e-3, "umma f32 relu")
    assert float((C[:, N:] - 7.0).abs().max() if Np > N else 0.0) == 0.0          # nothing written out of bounds
    # bf16 out, accumulate
    C0 = rnd((M, Np), dev, bf, 4)
    C1 = C0.clone()
    cb.umma_gemm(Av, a_mn, Bv, b_mn, C1[:, :N], None, M, N, K, GEMM_ACCUM)
    close(C1[:, :N], ref + C0[:, :N].float(), bf, what="umma bf16 accum")
    # split-K into fp32
    C2 = torch.zeros(M, Np, device=dev)
    cb.umma_gemm(Av, a_mn, Bv, b_mn, C2[:, :N], None, M, N, K, GEMM_SPLITK, 3)
    close(C2[:, :N], ref, torch.float32, 2e-3, "umma split-K")


@pytest.mark.parametrize("M,N,K", [(1056, 2048, 512), (4096, 512, 2048), (200, 96, 72), (130, 64, 264)])
def test_umma_gemm_fused_epilogues(dev, M, N, K):
    """masr_umma_gemm_ex extras: tensor-core row sums (bias gradient of a wgrad GEMM), ReLU+dropout fused into a
    forward GEMM (bit-identical to GEMM followed by masr_dropout), and the fused ReLU/dropout backward mask."""
    from metaasr_crossaccent_b200.ops import CudaBackend, GEMM_ACCUM, GEMM_RELU, GEMM_SPLITK
    cb = CudaBackend(dev, torch.bfloat16, gemm="umma")
    bf = torch.bfloat16
    x, w, bias = rnd((M, K), dev, bf, 1), rnd((N, K), dev, bf, 2, 0.1), rnd((N,), dev, torch.float32, 3)
    # forward: relu + dropout in the epilogue == separate dropout kernel on the same (seed, site)
    y1, y2 = torch.empty(M, N, device=dev, dtype=bf), torch.empty(M, N, device=dev, dtype=bf)
    cb.linear_fwd(x, w, bias, y1, relu=True, dropout=(0.1, 77, 5))
    cb.linear_fwd(x, w, bias, y2, relu=True)
    cb.dropout(y2, 0.1, 77, 5)
    assert torch.equal(y1 == 0, y2 == 0)         # same mask; values differ only by the rounding order
    close(y1, y2, bf, what="fused relu+dropout forward")
    keep = float((y1 != 0).float().mean()) / max(float((torch.relu(x.float() @ w.float().t() + bias) > 0).float().mean()), 1e-9)
    assert abs(keep - 0.9) < 0.02
    # backward of relu + dropout fused into the dgrad epilogue: dx = (f > 0) ? dgrad / (1 - p) : 0
    dy = rnd((M, K), dev, bf, 4)        # gradient wrt a [M, K]-shaped output of a second linear with weight w2 [K, N]
    w2 = rnd((K, N), dev, bf, 5, 0.1)
    g1 = torch.full((M, N), 3.0, device=dev, dtype=bf)
    cb.linear_dgrad(dy, w2, g1, relu_drop_mask=y1, p=0.1)
    ref = (dy.float() @ w2.float()) * (y1.float() > 0) / 0.9
    close(g1, ref, bf, what="fused relu/dropout backward")
    # wgrad with fused bias gradient (row sums of dy^T), with and without split-K
    dyo = rnd((M, N), dev, bf, 6)
    for sk in (1, 3):
        dw, db = torch.zeros(N, K, device=dev), rnd((N,), dev, torch.float32, 7)
        db0 = db.clone()
        cb.umma_gemm(dyo, 1, x, 1, dw, None, N, K, M, GEMM_SPLITK if sk > 1 else GEMM_ACCUM, sk, rowsum=db)
        close(dw, dyo.float().t() @ x.float(), torch.float32, 2e-3, "wgrad")
        close(db - db0, dyo.float().sum(0), torch.float32, 2e-3, "fused bias gradient")


@pytest.mark.parametrize("Bsz,L,H,N", [(32, 33, 8, 512), (4, 128, 8, 512), (3, 17, 2, 72)])
def test_umma_gemm_rowdot_epilogue(dev, Bsz, L, H, N):
    """Out-projection dgrad with the attention backward's D = rowsum(dO . O) per head produced by the GEMM epilogue
    (masr_gemm_epilogue.dot_*): dx unchanged, D[b,h,q] equals the separate reduction over the head's 64 columns."""
    from metaasr_crossaccent_b200.ops import CudaBackend
    cb = CudaBackend(dev, torch.bfloat16, gemm="umma")
    bf = torch.bfloat16
    M, K = Bsz * L, H * 64
    dy, w, o = rnd((M, N), dev, bf, 1), rnd((N, K), dev, bf, 2, 0.1), rnd((M, K), dev, bf, 3)
    dx, dx0 = torch.empty(M, K, device=dev, dtype=bf), torch.empty(M, K, device=dev, dtype=bf)
    dsum = torch.full((Bsz * H * L,), 9.0, device=dev)
    ready = cb.linear_dgrad(dy, w, dx, rowdot=(o, dsum, L, H))
    assert ready or N % 64 != 0        # the model's shapes must take the fused path
    cb.linear_dgrad(dy, w, dx0)
    assert torch.equal(dx, dx0)
    if not ready:
        return
    ref = ((dy.float() @ w.float()) * o.float()).view(Bsz, L, H, 64).sum(-1).permute(0, 2, 1).reshape(-1)
    close(dsum, ref, torch.float32, 5e-3, "row-dot epilogue")


@pytest.mark.parametrize("M,N,K", [(4096, 2048, 512), (4096, 512, 2048), (1000, 367, 520), (512, 2560, 4096), (300, 136, 200),
                                   (2304, 1536, 512)])
@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (0, 1), (1, 1)])
@pytest.mark.parametrize("bn", [128, 256])
@pytest.mark.parametrize("legacy_epilogue", [False, True])
def test_umma_gemm_pair(dev, M, N, K, a_mn, b_mn, bn, legacy_epilogue, request):
    """Persistent CTA-pair GEMM (tcgen05.mma.cta_group::2, 256 x bn tiles, double-buffered TMEM accumulator) against an
    fp32 matmul of the same bf16 operands: forward / dgrad / wgrad operand majors, bf16 and fp32 outputs, bias / ReLU /
    accumulate / split-K, ragged M / N / K (TMA zero fill, partial tiles), several tiles per pair (M = 4096: 128 items)."""
    from metaasr_crossaccent_b200.ops import CudaBackend, GEMM_ACCUM, GEMM_RELU, GEMM_SPLITK
    cb = CudaBackend(dev, torch.bfloat16, gemm="umma")
    # both epilogues: bulk tensor stores / reduce-adds (epilogue_tma.cuh, default) and the staged one (epilogue.cuh, which
    # also serves outputs a tensor map cannot describe)
    cb.lib.masr_gemm_set_pair_mode(3 if legacy_epilogue else 1)
    request.addfinalizer(lambda: cb.lib.masr_gemm_set_pair_mode(1))
    Mp, Np, Kp = (M + 7) // 8 * 8, (N + 7) // 8 * 8, (K + 7) // 8 * 8
    bf = torch.bfloat16
    A = rnd((Kp, Mp) if a_mn else (Mp, Kp), dev, bf, 1)
    B = rnd((Kp, Np) if b_mn else (Np, Kp), dev, bf, 2, 0.1)
    Af = (A[:K, :M].t() if a_mn else A[:M, :K]).float()
    Bf = (B[:K, :N].t() if b_mn else B[:N, :K]).float()
    bias = rnd((N,), dev, torch.float32, 3)
    ref = Af @ Bf.t()
    Av = A[:K] if a_mn else A[:M]
    Bv = B[:K] if b_mn else B[:N]
    C = torch.full((M, Np), 7.0, device=dev)
    cb.umma_gemm_pair(Av, a_mn, Bv, b_mn, C[:, :N], bias, M, N, K, GEMM_RELU, bn=bn)
    close(C[:, :N], torch.relu(ref + bias), torch.float32, 2e-3, "pair f32 relu")
    assert float((C[:, N:] - 7.0).abs().max() if Np > N else 0.0) == 0.0
    C0 = rnd((M, Np), dev, bf, 4)
    C1 = C0.clone()
    cb.umma_gemm_pair(Av, a_mn, Bv, b_mn, C1[:, :N], None, M, N, K, GEMM_ACCUM, bn=bn)
    close(C1[:, :N], ref + C0[:, :N].float(), bf, what="pair bf16 accum")
    Cb = torch.empty(M, Np, device=dev, dtype=bf)
    cb.umma_gemm_pair(Av, a_mn, Bv, b_mn, Cb[:, :N], bias, M, N, K, 0, bn=bn)
    close(Cb[:, :N], ref + bias, bf, what="pair bf16 store")
    for sk in (0, 3):
        C2 = torch.zeros(M, Np, device=dev)
        cb.umma_gemm_pair(Av, a_mn, Bv, b_mn, C2[:, :N], None, M, N, K, GEMM_SPLITK, splitk=sk, bn=bn)
        close(C2[:, :N], ref, torch.float32, 2e-3, f"pair split-K {sk}")
    C4 = rnd((M, Np), dev, torch.float32, 5)
    C40 = C4.clone()
    cb.umma_gemm_pair(Av, a_mn, Bv, b_mn, C4[:, :N], bias, M, N, K, GEMM_ACCUM, bn=bn)
    close(C4[:, :N], ref + bias + C40[:, :N], torch.float32, 2e-3, "pair f32 accumulate")
    # repeated launches reuse nothing stale (barrier phases, TMEM stages): same result twice in a row
    C3 = torch.empty(M, Np, device=dev, dtype=bf)
    cb.umma_gemm_pair(Av, a_mn, Bv, b_mn, C3[:, :N], bias, M, N, K, 0, bn=bn)
    assert torch.equal(C3[:, :N], Cb[:, :N])


@pytest.mark.parametrize("M,N,K", [(4096, 2048, 512), (1056, 2048, 512)])
def test_umma_gemm_pair_fused_epilogues(dev, M, N, K):
    """The fused extras on the pair kernel: ReLU + dropout forward (same mask as masr_dropout), ReLU/dropout backward
    mask in a dgrad, tensor-core row sums (bias gradient) of a split-K wgrad, per-head row dots of an out-projection dgrad."""
    from metaasr_crossaccent_b200.ops import CudaBackend, GEMM_ACCUM, GEMM_RELU, GEMM_SPLITK
    cb = CudaBackend(dev, torch.bfloat16, gemm="umma")
    bf = torch.bfloat16
    x, w, bias = rnd((M, K), dev, bf, 1), rnd((N, K), dev, bf, 2, 0.1), rnd((N,), dev, torch.float32, 3)
    y1, y2 = torch.empty(M, N, device=dev, dtype=bf), torch.empty(M, N, device=dev, dtype=bf)
    cb.umma_gemm_pair(x, 0, w, 0, y1, bias, M, N, K, GEMM_RELU, p_drop=0.1, seed=77, site=5)
    cb.umma_gemm_pair(x, 0, w, 0, y2, bias, M, N, K, GEMM_RELU)
    cb.dropout(y2, 0.1, 77, 5)
    assert torch.equal(y1 == 0, y2 == 0)
    close(y1, y2, bf, what="pair fused relu+dropout forward")
    dy = rnd((M, K), dev, bf, 4)
    w2 = rnd((K, N), dev, bf, 5, 0.1)
    g1 = torch.full((M, N), 3.0, device=dev, dtype=bf)
    cb.umma_gemm_pair(dy, 0, w2, 1, g1, None, M, N, K, 0, mask=y1, mask_scale=1.0 / 0.9)
    close(g1, (dy.float() @ w2.float()) * (y1.float() > 0) / 0.9, bf, what="pair fused relu/dropout backward")
    dyo = rnd((M, N), dev, bf, 6)
    for sk in (1, 0):
        dw, db = torch.zeros(N, K, device=dev), rnd((N,), dev, torch.float32, 7)
        db0 = db.clone()
        cb.umma_gemm_pair(dyo, 1, x, 1, dw, None, N, K, M, GEMM_SPLITK if sk != 1 else GEMM_ACCUM, splitk=sk, rowsum=db)
        close(dw, dyo.float().t() @ x.float(), torch.float32, 2e-3, "pair wgrad")
        close(db - db0, dyo.float().sum(0), torch.float32, 2e-3, "pair fused bias gradient")
    # row dots: out-projection dgrad of an attention block, 8 heads of 64
    H, L = 8, 33 if M % 33 == 0 else 128
    o = rnd((M, 512), dev, bf, 8)
    wo = rnd((N, 512), dev, bf, 9, 0.1)
    dx = torch.empty(M, 512, device=dev, dtype=bf)
    dsum = torch.full((M // L * H * L,), 9.0, device=dev)
    cb.umma_gemm_pair(dyo, 0, wo, 1, dx, None, M, 512, N, 0, rowdot=(o, dsum, L, H))
    refdx = dyo.float() @ wo.float()
    close(dx, refdx, bf, what="pair dgrad with row dots")
    ref = (refdx * o.float()).view(M // L, L, H, 64).sum(-1).permute(0, 2, 1).reshape(-1)
    close(dsum, ref, torch.float32, 5e-3, "pair row-dot epilogue")
