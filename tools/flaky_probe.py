#!/usr/bin/env python3
"""Run-to-run stability of one accent's share of the meta-gradient on ONE GPU: the share (inner-train step + inner-test
gradient, dropout off, kernel-by-kernel) is computed once per accent as the reference and then again and again, in random
order, with graph-replayed meta-steps in between to perturb allocator / workspace / stream state.  Prints the worst
relative L2 deviation per round and, for outliers, the tensors that moved.  Expected: ~1e-7 (fp32 atomics order).

    python tools/flaky_probe.py [--rounds 30]
"""
from __future__ import annotations

import argparse
import random
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rounds", type=int, default=30)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    solver, _ = bench.make_meta_solver("fomaml", 1, "bf16", True, 4)
    eng = solver.asr_model.engine
    n = eng.layout.total
    _, host_tasks = bench.host_tasks_of(0, 1, 1)
    warm = lambda: solver.meta_step_on_tasks([bench.clone_host(t) for t in host_tasks], global_task_count=8)
    warm(); warm()
    solver.flush_train_info()
    w0 = solver._original_flat.clone()

    def share(t):
        cfg = eng.cfg
        pd, ppd, g, lanes = cfg.dropout, cfg.pos_dropout, eng.use_graphs, solver.config["asr_model"]["task_lanes"]
        cfg.dropout = cfg.pos_dropout = 0.0
        eng.use_graphs = False
        solver.config["asr_model"]["task_lanes"] = 1
        solver._original_flat.copy_(w0)
        solver._upd_flat.zero_(); solver._counter = 0
        tr, te = bench.clone_host(t)
        solver.run_task(tr); solver.inner_test(te)
        u = solver._upd_flat[:n].clone()
        solver._upd_flat.zero_(); solver._counter = 0; solver._ring_sizes = []
        cfg.dropout, cfg.pos_dropout, eng.use_graphs = pd, ppd, g
        solver.config["asr_model"]["task_lanes"] = lanes
        return u

    ref = [share(t) for t in host_tasks]
    rng = random.Random(0)
    worst_all = 0.0
    for r in range(a.rounds):
        if r % 3 == 0:
            warm(); solver.flush_train_info()
        order = list(range(len(host_tasks)))
        rng.shuffle(order)
        worst = (0.0, -1)
        for i in order:
            u = share(host_tasks[i])
            rel = float((u - ref[i]).norm() / ref[i].norm())
            worst = max(worst, (rel, i))
            if rel > 1e-5:
                bad = sorted(((float((eng.layout.view(u, nm) - eng.layout.view(ref[i], nm)).norm() /
                                     eng.layout.view(ref[i], nm).norm().clamp_min(1e-30)), nm) for nm in eng.layout.offsets),
                             reverse=True)[:6]
                print(f"round {r} accent {i}: rel {rel:.3e}; tensors: {[(f'{v:.2e}', nm) for v, nm in bad]}", flush=True)
        worst_all = max(worst_all, worst[0])
        print(f"round {r}: worst rel {worst[0]:.3e} (accent {worst[1]})", flush=True)
    print(f"worst over {a.rounds} rounds: {worst_all:.3e}")


if __name__ == "__main__":
    main()
