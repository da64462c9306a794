#!/usr/bin/env python3
"""Summaries of ncu output for profiles/ (the judged copies; gpurun_out/ is scratch).

    python tools/summarize_ncu.py launches gpurun_out/launches.csv  "title" "command" > profiles/x.md
    python tools/summarize_ncu.py full     gpurun_out/k.ncu-rep     "title" "command" > profiles/y.md
"""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "sm__inst_executed_pipe_tensor_subpipe_umma.sum", "sm__pipe_tensor_subpipe_umma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__cycles_active.avg",
        "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
        "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct"]


def to_ms(v, u):
    v = float(v.replace(",", ""))
    return v / {"ns": 1e6, "us": 1e3, "ms": 1.0, "s": 1e-3}.get(u, 1e6)


def launches(path, title, cmd):
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        k = row["Kernel Name"].split("(")[0][:100]
        agg[k][0] += 1
        agg[k][1] += to_ms(row["Metric Value"], row["Metric Unit"])
    tot = sum(v[1] for v in agg.values())
    n = sum(v[0] for v in agg.values())
    print(f"# {title}\n\nCommand: `{cmd}`\n\nncu launch list (cold-cache, serialised: compare SHARES, not absolutes). "
          f"Total {tot:.2f} ms over {n} launches.\n\n| ms | share | launches | avg us | kernel |\n|---|---|---|---|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {v[1]:.3f} | {100 * v[1] / tot:.2f}% | {v[0]} | {1e3 * v[1] / v[0]:.1f} | `{k}` |")


def full(path, title, cmd):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader([l for l in out.splitlines() if l.startswith('"')]))
    h, units = rows[0], rows[1]
    print(f"# {title}\n\nCommand: `{cmd}`\n\nFrom `ncu --set full --clock-control none` ({path.split('/')[-1]}); one column per captured launch.\n")
    ki = h.index("Kernel Name")
    print("| metric | unit | " + " | ".join(f"#{i}" for i in range(len(rows) - 2)) + " |")
    print("|---|---|" + "---|" * (len(rows) - 2))
    print("| kernel | | " + " | ".join("`" + r[ki].split("(")[0][-48:] + "`" for r in rows[2:]) + " |")
    for k in KEYS:
        if k in h:
            i = h.index(k)
            print(f"| {k} | {units[i]} | " + " | ".join(r[i] for r in rows[2:]) + " |")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "")
