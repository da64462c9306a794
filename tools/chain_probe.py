#!/usr/bin/env python3
"""In-graph latency of the small kernels of the meta-step: N back-to-back (stream-ordered) launches of one
kernel are captured into a CUDA graph and replayed; replay time / N is what the kernel costs inside the
graph-replayed step (launch gap + prologue + execution + drain), which is what matters for the ~250 small
launches of a batch.

    python tools/chain_probe.py [--n 40] [--reps 5]
"""
from __future__ import annotations

import argparse
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from metaasr_crossaccent_b200.ops import CudaBackend  # noqa: E402


def chain(name, fn, n, reps, tab, flops=None):
    fn()
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(reps):
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3 / n)
    extra = f" {flops / best / 1e6:8.1f} TFLOP/s" if flops else ""
    tab.append(f"{name:56s} {best:8.2f} us/launch in graph{extra}")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=40)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    bf = torch.bfloat16
    be = CudaBackend(dev, bf, gemm="umma")
    r = lambda *s, dt=bf: torch.randn(*s, device=dev).to(dt)
    tab = []
    n, reps = args.n, args.reps
    for (M, N, K) in [(1056, 512, 512), (1056, 1536, 512), (1056, 2048, 512), (1056, 512, 2048), (1056, 367, 512),
                      (4096, 512, 512), (4096, 1536, 512), (4096, 2048, 512), (4096, 512, 2048), (4096, 512, 2560)]:
        x, w, b, y = r(M, K), r(N, K), r(N, dt=torch.float32), torch.empty(M, N, device=dev, dtype=bf)
        chain(f"linear_fwd   M{M} N{N} K{K}", lambda: be.linear_fwd(x, w, b, y), n, reps, tab, 2.0 * M * N * K)
        dy, dx = r(M, N), torch.empty(M, K, device=dev, dtype=bf)
        chain(f"linear_dgrad M{M} N{N} K{K}", lambda: be.linear_dgrad(dy, w, dx), n, reps, tab, 2.0 * M * N * K)
        dw, db = torch.zeros(N, K, device=dev), torch.zeros(N, device=dev)
        chain(f"linear_wgrad M{M} N{N} K{K} (no colsum)", lambda: be.linear_wgrad(x, dy, dw, None), n, reps, tab, 2.0 * M * N * K)
        chain(f"colsum       M{M} N{N}", lambda: be.colsum_add(dy, db), n, reps, tab)
    for (B, H, Lq, Lk, causal, kl) in [(32, 8, 128, 128, False, True), (32, 8, 33, 33, True, False), (32, 8, 33, 128, False, True)]:
        d = H * 64
        q, k, v = r(B * Lq, d), r(B * Lk, d), r(B * Lk, d)
        o, lse = torch.empty(B * Lq, d, device=dev, dtype=bf), torch.empty(B * H * Lq, device=dev)
        klens = torch.full((B,), Lk, dtype=torch.int64, device=dev) if kl else None
        fl = 4.0 * B * H * Lq * Lk * 64
        chain(f"attn_fwd B{B} H{H} Lq{Lq} Lk{Lk}", lambda: be.attn_fwd(q, k, v, o, lse, B, H, Lq, Lk, klens, causal, 0.1, 1, 1),
              n, reps, tab, fl)
        do, dq, dk, dv = r(B * Lq, d), torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
        ds = torch.empty(B * H * Lq, device=dev)
        chain(f"attn_bwd B{B} H{H} Lq{Lq} Lk{Lk} (2 kernels)",
              lambda: be.attn_bwd(q, k, v, o, do, lse, ds, dq, dk, dv, B, H, Lq, Lk, klens, causal, 0.1, 1, 1), n, reps, tab, 2.5 * fl)
    for rows in (4096, 1056):
        x, res, y = r(rows, 512), r(rows, 512), torch.empty(rows, 512, device=dev, dtype=bf)
        g, b = torch.ones(512, device=dev), torch.zeros(512, device=dev)
        m, rs = torch.empty(rows, device=dev), torch.empty(rows, device=dev)
        chain(f"add_layernorm_fwd rows{rows}", lambda: be.add_layernorm_fwd(x, res, g, b, y, m, rs, 0.1, 1, 1), n, reps, tab)
        ds, dx, dg, dbt = torch.empty_like(x), torch.empty_like(x), torch.zeros(512, device=dev), torch.zeros(512, device=dev)
        chain(f"add_layernorm_bwd rows{rows}", lambda: be.add_layernorm_bwd(y, x, m, rs, g, ds, False, dx, dg, dbt, 0.1, 1, 1), n, reps, tab)
        f1 = r(rows, 2048)
        chain(f"dropout rows{rows} x2048", lambda: be.dropout(f1, 0.1, 1, 2), n, reps, tab)
        chain(f"relu_bwd rows{rows} x2048", lambda: be.relu_bwd(f1, f1), n, reps, tab)
    print("\n".join(tab))


if __name__ == "__main__":
    main()
