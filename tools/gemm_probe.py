#!/usr/bin/env python3
"""Times the dense tcgen05 GEMM shapes of one hkust batch (B=32, T=512, L=32) on both kernels: the one-CTA-per-tile
kernel (gemm_umma.cu) and the persistent CTA-pair kernel (gemm_pair_umma.cu).  CUDA events, 20 launches after 3 warm-ups,
L2 flushed between launches by a 256 MB memset.  Usage: python tools/gemm_probe.py [--md out.md]"""
import argparse
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from metaasr_crossaccent_b200.ops import CudaBackend, GEMM_ACCUM, GEMM_SPLITK  # noqa: E402

SHAPES = [  # (what, M, N, K, a_mn, b_mn, flags)
    ("enc qkv fwd", 4096, 1536, 512, 0, 0, 0), ("enc out fwd", 4096, 512, 512, 0, 0, 0),
    ("enc ff1 fwd", 4096, 2048, 512, 0, 0, 0), ("enc ff2 fwd", 4096, 512, 2048, 0, 0, 0),
    ("vgg2enc fwd", 4096, 512, 2560, 0, 0, 0), ("dec kv fwd", 4096, 1024, 512, 0, 0, 0),
    ("enc qkv dgrad", 4096, 512, 1536, 0, 1, 0), ("enc ff1 dgrad", 4096, 512, 2048, 0, 1, 0),
    ("enc ff2 dgrad", 4096, 2048, 512, 0, 1, 0), ("vgg2enc dgrad", 4096, 2560, 512, 0, 1, 0),
    ("enc qkv wgrad", 1536, 512, 4096, 1, 1, GEMM_SPLITK), ("enc ff1 wgrad", 2048, 512, 4096, 1, 1, GEMM_SPLITK),
    ("enc ff2 wgrad", 512, 2048, 4096, 1, 1, GEMM_SPLITK), ("vgg2enc wgrad", 512, 2560, 4096, 1, 1, GEMM_SPLITK),
    ("dec ff1 fwd", 1056, 2048, 512, 0, 0, 0), ("dec ff2 fwd", 1056, 512, 2048, 0, 0, 0), ("dec out fwd", 1056, 512, 512, 0, 0, 0),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--md", default=None)
    ap.add_argument("--only", default=None, help="substring of the GEMM name")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--warm", action="store_true", help="no L2 flush: 50 launches back to back inside one event pair")
    ap.add_argument("--graph", action="store_true", help="no L2 flush, no host launch cost: 20 launches captured in a CUDA graph")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    cb = CudaBackend(dev, torch.bfloat16, gemm="umma")
    flush = torch.empty(64 << 20, dtype=torch.float32, device=dev)
    rows = []
    for what, M, N, K, a_mn, b_mn, flags in SHAPES:
        if args.only and args.only not in what:
            continue
        A = torch.randn((K, M) if a_mn else (M, K), device=dev).to(torch.bfloat16)
        B = torch.randn((K, N) if b_mn else (N, K), device=dev).to(torch.bfloat16)
        C = torch.zeros(M, N, device=dev, dtype=torch.float32 if flags & GEMM_SPLITK else torch.bfloat16)
        res = {}
        for mode, fn in (("tile_legacy", lambda: (cb.lib.masr_gemm_set_pair_mode(2), cb.umma_gemm(A, a_mn, B, b_mn, C, None, M, N, K, flags, cb._wgrad_splitk(M, N, K) if flags else 1))),
                         ("tile", lambda: (cb.lib.masr_gemm_set_pair_mode(0), cb.umma_gemm(A, a_mn, B, b_mn, C, None, M, N, K, flags, cb._wgrad_splitk(M, N, K) if flags else 1))),
                         ("pair128", lambda: cb.umma_gemm_pair(A, a_mn, B, b_mn, C, None, M, N, K, flags, 0, 128)),
                         ("pair256", lambda: cb.umma_gemm_pair(A, a_mn, B, b_mn, C, None, M, N, K, flags, 0, 256)),
                         ("pair", lambda: cb.umma_gemm_pair(A, a_mn, B, b_mn, C, None, M, N, K, flags, 0, 0))):
            for _ in range(3):
                fn()
            tot = 0.0
            if args.graph:
                torch.cuda.synchronize()
                st = torch.cuda.Stream()
                with torch.cuda.stream(st):
                    fn()
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=st):
                        for _ in range(20):
                            fn()
                    g.replay()
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(5):
                        g.replay()
                    e1.record()
                    torch.cuda.synchronize()
                res[mode] = e0.elapsed_time(e1) / 100 * 1e3
                continue
            if args.warm:
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(50):
                    fn()
                e1.record()
                torch.cuda.synchronize()
                res[mode] = e0.elapsed_time(e1) / 50 * 1e3
                continue
            for _ in range(args.iters):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(); e1.record()
                torch.cuda.synchronize()
                tot += e0.elapsed_time(e1)
            res[mode] = tot / args.iters * 1e3
        cb.lib.masr_gemm_set_pair_mode(1)
        gf = 2.0 * M * N * K
        rows.append((what, M, N, K, res))
        print(f"{what:16s} M={M:5d} N={N:5d} K={K:5d}  " + "  ".join(f"{m} {us:6.1f} us {gf / us / 1e6:6.0f} TF/s" for m, us in res.items()), flush=True)
    if args.md:
        with open(args.md, "w") as f:
            f.write("| GEMM | M | N | K | tile, staged epilogue us | tile us | pair128 us | pair256 us | pair(auto) us | auto TFLOP/s |\n|---|---|---|---|---|---|---|---|---|---|\n")
            for what, M, N, K, r in rows:
                f.write(f"| {what} | {M} | {N} | {K} | {r['tile_legacy']:.1f} | {r['tile']:.1f} | {r['pair128']:.1f} | {r['pair256']:.1f} | {r['pair']:.1f} | {2.0 * M * N * K / r['pair'] / 1e6:.0f} |\n")


if __name__ == "__main__":
    main()
