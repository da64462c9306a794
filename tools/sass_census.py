#!/usr/bin/env python3
"""Instruction census of the shipped library: `cuobjdump -sass libmetaasr_b200.so`, per kernel the counts of the SASS
mnemonics that prove the Blackwell paths (tcgen05.mma = UTCHMMA / UTCQMMA..., TMEM loads = LDTM, TMEM alloc = UTCALLOC...,
TMA = UTMALDG / UTMASTG / UTMAREDG / UBLKCP, mbarrier = SYNCS, tcgen05.commit = UTCBAR) next to the CUDA-core work
(FFMA, MUFU; HMMA = warp-level mma.sync: only the short-query attention kernels of attn_small.cu, on purpose).  Usage: python tools/sass_census.py [out.md]"""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
LIB = ROOT / "metaasr_crossaccent_b200" / "libmetaasr_b200.so"
COLS = ["UTCHMMA", "UTCBAR", "LDTM", "UTCATOMSWS", "UTMALDG", "UTMASTG", "UTMAREDG", "SYNCS", "ELECT", "UCGABAR", "HMMA", "FFMA",
        "MUFU", "SHFL", "REDUX", "LDG", "STG", "RED", "LDS", "STS"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m:
            op = m.group(1)
            kernels[cur][op] += 1
            kernels[cur]["_total"] += 1
    demangled = subprocess.run(["cu++filt"] + list(kernels), capture_output=True, text=True).stdout.splitlines()
    rows = []
    for (name, cnt), dn in zip(kernels.items(), demangled):
        dn = re.sub(r"\((int|bool|unsigned int)\)", "", dn).replace("(anonymous namespace)::", "")
        short = re.sub(r"\(.*", "", dn).replace("void ", "").replace("masr::", "")
        rows.append((short, cnt))
    rows.sort(key=lambda r: (-r[1]["UTCHMMA"], r[0]))
    lines = ["# SASS instruction census of libmetaasr_b200.so (sm_100a)", "",
             f"`cuobjdump -sass {LIB.relative_to(ROOT)}` -> static instruction counts per kernel ({len(rows)} kernels).",
             "UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, LDTM = tcgen05.ld, UTCATOMSWS = tcgen05.alloc / dealloc, UTMALDG /",
             "UTMASTG / UTMAREDG = TMA tensor load / store / reduce-add, SYNCS = mbarrier ops, ELECT = elect.sync (converged-warp",
             "issue), UCGABAR = cluster barrier (CTA-pair GEMM), HMMA = warp-level mma.sync (only attn_small_*: the 33-row decoder attention problems, latency-bound, DESIGN.md 2c').", "",
             "| kernel | total | " + " | ".join(COLS) + " |", "|---|---|" + "---|" * len(COLS)]
    for short, cnt in rows:
        lines.append(f"| `{short}` | {cnt['_total']} | " + " | ".join(str(sum(v for k, v in cnt.items() if k.startswith(c))) for c in COLS) + " |")
    tot_tc = sum(1 for _, c in rows if c["UTCHMMA"] > 0)
    lines += ["", f"{tot_tc} kernels issue tcgen05.mma; {sum(1 for _, c in rows if any(k.startswith('HMMA') for k in c))} use warp-level HMMA (attn_small_fwd / attn_small_bwd)."]
    text = "\n".join(lines) + "\n"
    if len(sys.argv) > 1:
        Path(sys.argv[1]).write_text(text)
    else:
        print(text)


if __name__ == "__main__":
    main()
