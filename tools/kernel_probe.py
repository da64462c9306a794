#!/usr/bin/env python3
"""Launch each hot kernel of the meta-step a few times at the BASELINE shapes (B=32, T=512, L=32), so that
`ncu` can profile them in isolation without replaying a whole meta-step:

    python tools/kernel_probe.py [--reps 3] [--only gemm|conv|attn|elem|mt|ctc]
    ncu --set full --clock-control none --import-source on -o gpurun_out/probe python tools/kernel_probe.py --reps 1

Also prints CUDA-event timings (mean over reps after one warm-up launch) as a quick per-kernel table.
"""
from __future__ import annotations

import argparse
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from metaasr_crossaccent_b200.ops import CudaBackend, GEMM_ACCUM, GEMM_RELU, GEMM_SPLITK  # noqa: E402


def timeit(name, fn, reps, table, flops=None, bytes_=None):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    inner = 10 if reps > 1 else 1          # back-to-back launches hide the host launch latency
    for _ in range(reps):
        e0.record()
        for _ in range(inner):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3 / inner)
    us = sum(ts) / len(ts)
    extra = ""
    if flops:
        extra += f" {flops / us / 1e6:8.1f} TFLOP/s"
    if bytes_:
        extra += f" {bytes_ / us / 1e3:8.1f} GB/s"
    table.append(f"{name:52s} {us:9.1f} us{extra}")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--only", default=None)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    bf = torch.bfloat16
    be = CudaBackend(dev, bf, gemm="umma")
    r = lambda *s, dt=bf: torch.randn(*s, device=dev).to(dt)
    tab = []
    want = lambda k: args.only is None or args.only == k

    if want("gemm"):
        for (M, N, K) in [(1056, 512, 512), (1056, 1536, 512), (1056, 2048, 512), (1056, 512, 2048),
                          (4096, 1536, 512), (4096, 2048, 512), (4096, 512, 2048), (4096, 512, 2560)]:
            x, w, b, y = r(M, K), r(N, K), r(N, dt=torch.float32), torch.empty(M, N, device=dev, dtype=bf)
            timeit(f"linear_fwd   M{M} N{N} K{K}", lambda: be.linear_fwd(x, w, b, y), args.reps, tab, 2.0 * M * N * K)
            dy, dx = r(M, N), torch.empty(M, K, device=dev, dtype=bf)
            timeit(f"linear_dgrad M{M} N{N} K{K}", lambda: be.linear_dgrad(dy, w, dx), args.reps, tab, 2.0 * M * N * K)
            dw, db = torch.zeros(N, K, device=dev), torch.zeros(N, device=dev)
            timeit(f"linear_wgrad M{M} N{N} K{K} (+colsum)", lambda: be.linear_wgrad(x, dy, dw, db), args.reps, tab, 2.0 * M * N * K)
    if want("conv"):
        for (B, H, W, Ci, Co) in [(32, 512, 83, 64, 64), (32, 256, 41, 64, 128), (32, 256, 41, 128, 128)]:
            x, w, bias = r(B, H, W, Ci), torch.randn(Co, Ci, 3, 3, device=dev) * 0.05, r(Co, dt=torch.float32)
            wp = torch.empty(Co, 9 * Ci, device=dev, dtype=bf)
            be.conv_w_prep(w, wp)
            y, dy = torch.empty(B, H, W, Co, device=dev, dtype=bf), r(B, H, W, Co)
            fl = 2.0 * B * H * W * Co * 9 * Ci
            timeit(f"conv3x3_fwd   {B}x{H}x{W} {Ci}->{Co}", lambda: be.conv3x3_fwd(x, wp, bias, y), args.reps, tab, fl)
            dx = torch.empty_like(x)
            wpt = torch.empty(Ci, 9 * Co, device=dev, dtype=bf)
            be.conv_w_prep_t(w, wpt)
            timeit(f"conv3x3_dgrad {B}x{H}x{W} {Ci}->{Co}", lambda: be.conv3x3_dgrad(dy, wp, dx, x, wpt=wpt), args.reps, tab, fl)
            dwp, db = torch.zeros(Co, 9 * Ci, device=dev), torch.zeros(Co, device=dev)
            timeit(f"conv3x3_wgrad {B}x{H}x{W} {Ci}->{Co} (+colsum)", lambda: be.conv3x3_wgrad(x, dy, dwp, db), args.reps, tab, fl)
        x1 = torch.randn(32, 512, 83, device=dev)
        w1, b1 = torch.randn(64, 1, 3, 3, device=dev), torch.randn(64, device=dev)
        a1 = torch.empty(32, 512, 83, 64, device=dev, dtype=bf)
        nb = a1.numel() * 2
        timeit("conv1_fwd 32x512x83 ->64", lambda: be.conv1_fwd(x1, w1, b1, a1), args.reps, tab, None, nb)
        dw1, db1 = torch.zeros(64, 1, 3, 3, device=dev), torch.zeros(64, device=dev)
        timeit("conv1_wgrad", lambda: be.conv1_wgrad(x1, a1, dw1, db1), args.reps, tab, None, nb)
        p1 = torch.empty(32, 256, 41, 64, device=dev, dtype=bf)
        timeit("maxpool_fwd 32x512x83x64", lambda: be.maxpool_fwd(a1, p1), args.reps, tab, None, nb + p1.numel() * 2)
        g1 = torch.empty_like(a1)
        timeit("maxpool_bwd 32x512x83x64", lambda: be.maxpool_bwd(a1, p1, g1, True), args.reps, tab, None, 2 * nb + p1.numel() * 2)
        code = torch.empty(p1.shape, device=dev, dtype=torch.uint8)
        timeit("maxpool_fwd + arg-max codes", lambda: be.maxpool_fwd(a1, p1, code=code), args.reps, tab, None, nb + p1.numel() * 3)
        timeit("maxpool_bwd from codes", lambda: be.maxpool_bwd(a1, p1, g1, True, code=code), args.reps, tab, None, nb + p1.numel() * 3)
    if want("attn"):
        for (B, H, Lq, Lk, causal, kl) in [(32, 8, 128, 128, False, True), (32, 8, 33, 33, True, False), (32, 8, 33, 128, False, True)]:
            d = H * 64
            q, k, v = r(B * Lq, d), r(B * Lk, d), r(B * Lk, d)
            o, lse = torch.empty(B * Lq, d, device=dev, dtype=bf), torch.empty(B * H * Lq, device=dev)
            klens = torch.full((B,), Lk, dtype=torch.int64, device=dev) if kl else None
            fl = 4.0 * B * H * Lq * Lk * 64
            timeit(f"attn_fwd B{B} H{H} Lq{Lq} Lk{Lk}", lambda: be.attn_fwd(q, k, v, o, lse, B, H, Lq, Lk, klens, causal, 0.1, 1, 1),
                   args.reps, tab, fl)
            do, dq, dk, dv = r(B * Lq, d), torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
            ds = torch.empty(B * H * Lq, device=dev)
            timeit(f"attn_bwd B{B} H{H} Lq{Lq} Lk{Lk}",
                   lambda: be.attn_bwd(q, k, v, o, do, lse, ds, dq, dk, dv, B, H, Lq, Lk, klens, causal, 0.1, 1, 1), args.reps, tab, 2.5 * fl)
    if want("elem"):
        for rows in (4096, 1056):
            x, res, y = r(rows, 512), r(rows, 512), torch.empty(rows, 512, device=dev, dtype=bf)
            g, b = torch.ones(512, device=dev), torch.zeros(512, device=dev)
            m, rs = torch.empty(rows, device=dev), torch.empty(rows, device=dev)
            timeit(f"add_layernorm_fwd rows{rows}", lambda: be.add_layernorm_fwd(x, res, g, b, y, m, rs, 0.1, 1, 1), args.reps, tab, None, rows * 512 * 2 * 4)
            ds, dx, dg, dbt = torch.empty_like(x), torch.empty_like(x), torch.zeros(512, device=dev), torch.zeros(512, device=dev)
            timeit(f"add_layernorm_bwd rows{rows}", lambda: be.add_layernorm_bwd(y, x, m, rs, g, ds, False, dx, dg, dbt, 0.1, 1, 1),
                   args.reps, tab, None, rows * 512 * 2 * 4)
            f1 = r(rows, 2048)
            timeit(f"dropout rows{rows} x2048", lambda: be.dropout(f1, 0.1, 1, 2), args.reps, tab, None, rows * 2048 * 4)
            cs = torch.zeros(2048, device=dev)
            timeit(f"colsum rows{rows} x2048", lambda: be.colsum_add(f1, cs), args.reps, tab, None, rows * 2048 * 2)
        lg, gold = torch.randn(1056, 367, device=dev), torch.randint(0, 367, (1056,), device=dev)
        st, am, dl = torch.zeros(4, dtype=torch.float64, device=dev), torch.empty(1056, dtype=torch.int64, device=dev), torch.empty(1056, 367, device=dev)
        timeit("ls_ce 1056x367", lambda: be.ls_ce(lg, gold, 0.2, 1 / 1056, st, am, dl), args.reps, tab, None, 1056 * 367 * 8)
    if want("mt"):
        n = 24_900_000 // 64 * 64
        p, g, buf, u = (torch.randn(n, device=dev) for _ in range(4))
        ss = torch.zeros(1, dtype=torch.float64, device=dev)
        timeit("mt_sumsq 24.9M", lambda: be.mt_sumsq(g, ss), args.reps, tab, None, n * 4)
        timeit("mt_clip_sgd 24.9M", lambda: be.mt_clip_sgd(p, g, buf, ss, 5.0, 1e-4, 0.9, True, False), args.reps, tab, None, n * 24)
        timeit("mt_accumulate 24.9M", lambda: be.mt_accumulate(u, g, ss, 5.0), args.reps, tab, None, n * 12)
        m, v = torch.zeros(n, device=dev), torch.zeros(n, device=dev)
        timeit("mt_adam 24.9M", lambda: be.mt_adam(p, m, v, u, 8.0, 1e-4, 0.9, 0.98, 1e-9, 0.1, 0.02), args.reps, tab, None, n * 28)
        pb = torch.empty(n, device=dev, dtype=bf)
        timeit("cast f32->bf16 24.9M", lambda: be.cast(p, pb), args.reps, tab, None, n * 6)
    if want("ctc"):
        # raw C-ABI call with device-resident arguments (the Python wrapper's host work would dominate)
        for (T, B, C, L) in [(128, 32, 367, 34), (128, 148, 367, 34), (128, 512, 367, 34), (128, 2048, 367, 34),
                             (375, 512, 367, 102), (750, 256, 367, 152)]:
            lg = torch.randn(T, B, C, device=dev)
            tg = torch.randint(1, C, (B * L,), device=dev)
            offs = torch.arange(B, device=dev, dtype=torch.int64) * L
            il = torch.full((B,), T, dtype=torch.int64, device=dev)
            tl = torch.full((B,), L, dtype=torch.int64, device=dev)
            nll, loss, grad = torch.empty(B, device=dev), torch.empty(1, device=dev), torch.empty_like(lg)
            wsb = be.lib.masr_ctc_workspace_bytes(T, B, C, L)
            ws = torch.empty(wsb // 4 + 1, device=dev) if wsb else None

            def run():
                rc = be.lib.masr_ctc_fwd_bwd(lg.data_ptr(), T, B, C, 0, tg.data_ptr(), offs.data_ptr(), il.data_ptr(),
                                             tl.data_ptr(), L, 0, 1, 1.0, nll.data_ptr(), loss.data_ptr(), grad.data_ptr(),
                                             ws.data_ptr() if ws is not None else None, wsb, be.stream)
                assert rc == 0
            timeit(f"ctc_fwd_bwd T{T} B{B} C{C} L{L}", run, args.reps, tab, None, T * B * C * 8)
            import ctypes
            be.lib.masr_ctc_debug_enable(1)
            run()
            buf = (ctypes.c_longlong * 6)()
            be.lib.masr_ctc_debug_read(buf)
            be.lib.masr_ctc_debug_enable(0)
            st = list(buf)
            # block-barrier kernels (v2, MASR_CTC_PIPE=0): emissions | recursions | gradient phase; pipelined kernel:
            # [1..2] = the alpha recursion running alongside the emission / gradient workers, [2..3] = their tail
            tab.append(f"    CTA0 cycles: setup {st[1]-st[0]}, [1..2] {st[2]-st[1]}, [2..3] {st[3]-st[2]}, [3..end] {st[5]-st[3]}")
    print("\n".join(tab))


if __name__ == "__main__":
    main()
