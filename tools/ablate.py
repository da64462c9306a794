#!/usr/bin/env python3
"""Timing ablation of the headline meta-step: the named backend entry points are replaced by no-ops (results are then
garbage -- this is a TIMING experiment only) and the step is timed as bench.py does.  The drop in ms/step is what the
entry point really costs inside the multi-lane, graph-replayed step (where HBM-bound and tensor-bound kernels of
different lanes overlap), which the per-launch times cannot tell.

    python tools/ablate.py [--lanes 4] [--steps 6] --skip maxpool_fwd,maxpool_bwd [--skip conv1_fwd,conv1_wgrad ...]
"""
from __future__ import annotations

import argparse
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def one(skip, lanes, steps):
    import torch
    import bench
    from metaasr_crossaccent_b200.ops import CudaBackend
    for name in skip:
        if name:
            assert hasattr(CudaBackend, name), name
            setattr(CudaBackend, name, lambda self, *a, **k: None)
    args = argparse.Namespace(no_graphs=False, lanes=lanes)
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    solver, _ = bench.make_meta_solver("fomaml", 1, "bf16", True, lanes)
    solver.asr_model.engine.be.strict_umma = False
    eng = solver.asr_model.engine
    _, host_tasks = bench.host_tasks_of(0, 1, 1)

    def prepared(task):
        tr, te = bench.clone_host(task)
        mk = lambda b: (b[0], (eng.to_device(eng.prepare_batch(*b[1])), None, [None] * bench.INNER_B, None))
        return [mk(b) for b in tr], mk(te)

    dev_tasks = [prepared(t) for t in host_tasks]
    torch.cuda.synchronize()
    fn = lambda: solver.meta_step_on_tasks(dev_tasks, global_task_count=bench.N_ACCENTS)
    ms, _, _, _ = bench.timed(fn, steps, 3, dev, 1, solver.backend)
    print(json.dumps({"skip": skip, "lanes": lanes, "ms_per_step": round(ms, 3)}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--skip", action="append", default=[])
    ap.add_argument("--lanes", type=int, default=4)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--child", default=None)
    a = ap.parse_args()
    if a.child is not None:
        one([s for s in a.child.split(",") if s], a.lanes, a.steps)
        return
    for s in [""] + a.skip:
        subprocess.run([sys.executable, __file__, "--child", s, "--lanes", str(a.lanes), "--steps", str(a.steps)], check=False)


if __name__ == "__main__":
    main()
