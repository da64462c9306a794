#!/usr/bin/env python3
"""bench.py -- meta-train frames/s of the FOMAML hot path (BASELINE.json metric / configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--dtype bf16|fp32]

A "step" is ONE FOMAML meta-step of the fometa-hkust network (d512/h8/ff2048/2e4d, C=367, label smoothing
0.2, dropout 0.1): 8 synthetic accents, meta_k = 1, inner batch 32 x 512 frames x 83-dim fbank+pitch, targets
of 32 unigram150 ids -> per accent one inner-train batch (fwd+bwd, clip, nesterov-SGD) and one inner-test
batch (fwd+bwd, clip, accumulate), then all-reduce + noam-Adam meta-update = 262 144 input frames per step.
Accents are partitioned over the ranks (strong scaling: the meta-batch of 8 accents is fixed).

  value  : frames/s with the step's batches already resident in HBM (CUDA events, max over ranks)
  e2e    : same metric through the public drop-in API (get_trainer / run_task / run_batch) from pinned HOST
           buffers: host->device copies of every batch and the device->host read of the losses are inside
           the timed region
  --impl reference : the reference's algorithm on the host CPU cores (oracle/port.py, the checker that is
           pinned to the live reference by tests/golden), one bounded sample per step.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

N_ACCENTS, META_K, INNER_B, T_FRAMES, L_TGT, IDIM = 8, 1, 32, 512, 32, 83
FRAMES_PER_STEP = N_ACCENTS * (META_K + 1) * INNER_B * T_FRAMES
WORKLOAD = ("FOMAML meta-step, fometa-hkust transformer (d512 h8 ff2048 2enc 4dec, C=367), 8 synthetic accents, "
            "meta_k 1, inner batch 32 x T512 x 83-dim fbank, L=32 unigram150 ids")


def hkust_config(dtype, gemm, dropout=0.1, graphs=True, lanes=1):
    am = {"idim": IDIM, "nheads": 8, "d_model": 512, "d_inner": 2048, "dropout": dropout, "tgt_share_weight": 1,
          "encoder": {"nlayers": 2}, "decoder": {"nlayers": 4}, "pos_dropout": dropout, "dtype": dtype, "gemm": gemm,
          "cuda_graphs": graphs, "task_lanes": lanes,
          "inner_optimizer_cls": "SGD", "inner_optimizer_opt": {"momentum": 0.9, "nesterov": True},
          "meta_opt_cls": "noam", "meta": {"optimizer_opt": {"k": 1.0, "warmup_steps": 25000}}}
    solver = {"setting": "fometa-transformer-hkust", "total_steps": 1000000, "label_smoothing": 0.2,
              "eval_ival": 5000, "log_ival": 20, "save_ival": 5000, "batch_size": 32}
    return {"asr_model": am, "solver": solver}


def synth_batch(gen, B=INNER_B, T=T_FRAMES, L=L_TGT, pin=False):
    """Profile P-eq of SURVEY 8(d): what the reference's 1-frame-bucket train loader yields."""
    x = torch.randn(B, T, IDIM, generator=gen)
    ilens = torch.full((B,), T, dtype=torch.int64)
    ys = [torch.randint(1, 366, (L,), generator=gen, dtype=torch.int64) for _ in range(B)]
    olens = torch.full((B,), L, dtype=torch.int64)
    if pin and torch.cuda.is_available():
        x = x.pin_memory()
    return x, ilens, ys, olens


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.lines:
            if ts < t0 - 0.05 or ts > t1 + 0.05:
                continue
            f = [v.strip() for v in line.split(",")]
            try:
                sm.append(float(f[1])); mx = float(f[2])
            except (ValueError, IndexError):
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ================================================================================================= ours
def run_ours(args):
    from metaasr_crossaccent_b200 import dist as D
    from metaasr_crossaccent_b200 import interfaces as I
    from metaasr_crossaccent_b200.trainer import get_trainer

    assert torch.cuda.is_available(), "bench.py measures the CUDA path; there is no CPU fallback"
    rank, world, local = D.init_from_env("nccl")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    gemm = "umma" if args.dtype == "bf16" else "simt"
    accents = ["af", "au", "ca", "en", "in", "ir", "nz", "us"]
    id2accent = {a: a for a in accents + ["hk"]}
    paras = argparse.Namespace(pretrain_accents=accents, num_pretrain=N_ACCENTS, tgt_accent="hk", runs=0, seed=531,
                               meta_k=META_K, meta_batch_size=N_ACCENTS, max_step=0, resume=False, algo="fomaml",
                               pretrain_suffix="bench", log_root=None)
    import random
    random.seed(531); torch.manual_seed(531)
    solver = get_trainer(I.FOMetaASRInterface, hkust_config(args.dtype, gemm, graphs=not args.no_graphs, lanes=args.lanes),
                         paras, id2accent)
    solver.set_model()
    eng, be = solver.asr_model.engine, solver.backend

    mine = D.partition_tasks(list(range(N_ACCENTS)), N_ACCENTS, rank, world)
    gen = torch.Generator().manual_seed(531 + rank)
    # synthetic pinned host batches: per owned accent, one inner-train and one inner-test batch
    host_tasks = [([(a, synth_batch(gen, pin=True)) for _ in range(META_K)], (a, synth_batch(gen, pin=True))) for a in mine]

    def clone_host(task):
        tr, te = task
        cl = lambda b: (b[0], (b[1][0], b[1][1].clone(), b[1][2], b[1][3].clone()))   # olens is mutated in place
        return [cl(b) for b in tr], cl(te)

    def prepared(task):
        tr, te = clone_host(task)
        mk = lambda b: (b[0], (eng.to_device(eng.prepare_batch(*b[1])), None, [None] * INNER_B, None))
        return [mk(b) for b in tr], mk(te)

    dev_tasks = [prepared(t) for t in host_tasks]            # inputs resident in HBM
    torch.cuda.synchronize()

    def step_resident():
        solver.meta_step_on_tasks(dev_tasks, global_task_count=N_ACCENTS)

    def step_e2e():
        solver.meta_step_on_tasks([clone_host(t) for t in host_tasks], global_task_count=N_ACCENTS)
        return solver.flush_train_info()                     # device->host read of the step's losses

    def timed(fn, steps, warmup, profile=False):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize(); D.barrier()
        be.launches = 0
        be.prof = {} if profile else None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize(); D.barrier()
        t1 = time.time()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
        prof, be.prof = be.prof, None
        return float(ms) / steps, be.launches // max(steps, 1), prof, (t0, t1)

    sampler = ClockSampler(local)
    sampler.start()
    graphs = eng.use_graphs
    ms_res, launches, prof, (t0, t1) = timed(step_resident, args.steps, args.warmup, profile=not graphs)
    clocks = sampler.stop(t0, t1)
    ms_e2e, _, _, _ = timed(step_e2e, args.steps, max(1, args.warmup // 2))
    if graphs:
        # graph replays bypass the host-side per-launch hooks: count launches and time the tensor-core kernels
        # (CUDA events on the launching stream) over extra meta-steps issued kernel by kernel, one accent at a
        # time (task_lanes = 1) so that no other lane's kernels share the GPU with the kernel being timed
        lanes_cfg = solver.config["asr_model"].get("task_lanes", 1)
        solver.config["asr_model"]["task_lanes"] = 1
        eng.use_graphs, ms_cfg, eng.multi_stream = False, eng.multi_stream, False     # no side stream either
        step_resident()
        _, launches, prof, _ = timed(step_resident, 2, 0, profile=True)
        eng.use_graphs, eng.multi_stream = True, ms_cfg
        solver.config["asr_model"]["task_lanes"] = lanes_cfg
        prof_steps = 2
    else:
        prof_steps = args.steps
    if args.profile and rank == 0:
        eng.use_graphs = False
        be.prof_ops = {}
        step_resident()
        torch.cuda.synchronize()
        po, be.prof_ops = be.prof_ops, None
        eng.use_graphs = graphs
        if prof:
            with open(args.profile + ".gemm", "w") as f:
                f.write("# tensor-core launches by shape over the timed region (CUDA events)\n")
                f.write("| total ms | calls | avg us | TFLOP/s | kind | M | N | K |\n|---|---|---|---|---|---|---|---|\n")
                rows = sorted(((sum(a.elapsed_time(b) for a, b in v), len(v), k) for k, v in prof.items()), reverse=True)
                for ms, n, (kind, M, N, K) in rows:
                    f.write(f"| {ms:.2f} | {n} | {1e3 * ms / n:.1f} | {2.0 * M * N * K * n / (ms * 1e-3) / 1e12:.1f} | {kind} | {M} | {N} | {K} |\n")
        rows = sorted(((sum(a.elapsed_time(b) for a, b in v), len(v), k) for k, v in po.items()), reverse=True)
        tot = sum(r[0] for r in rows)
        with open(args.profile, "w") as f:
            f.write(f"# per-entry-point CUDA-event times of ONE meta-step ({tot:.1f} ms summed; dtype {args.dtype})\n")
            f.write("| ms | share | calls | entry point |\n|---|---|---|---|\n")
            for ms, n, k in rows:
                f.write(f"| {ms:.2f} | {100 * ms / tot:.1f}% | {n} | {k} |\n")
    loss_info = solver.flush_train_info()

    # ---- roofline of the dominant kernel, timed live with CUDA events inside the timed region
    peaks = {}
    pk = ROOT / "MEASURED_PEAKS.json"
    if pk.exists():
        peaks = json.loads(pk.read_text())
    roof = None
    roof_all = []
    if prof:
        traffic = {}
        tj = ROOT / "profiles" / "r1_traffic.json"
        if tj.exists():
            traffic = json.loads(tj.read_text()).get("per_launch", {})
        peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
        rows = []
        for key, evs in prof.items():
            tot = sum(a.elapsed_time(b) for a, b in evs)
            rows.append((tot, key, len(evs)))
        rows.sort(reverse=True)
        # launches shorter than ~50 us are dominated by the host launch gap when issued kernel by kernel (they
        # take 4-15 us inside the replayed graph): the dominant kernel is chosen among the long launches
        long_rows = [r for r in rows if r[0] / r[2] >= 0.05] or rows
        rows = long_rows + [r for r in rows if r not in long_rows]
        for tot_ms, (kind, M, N, K), n in rows[:8]:
            ach = 2.0 * M * N * K / (tot_ms / n * 1e-3) / 1e12
            roof_all.append({"kernel": f"{kind} M={M} N={N} K={K}", "launches": n, "avg_launch_ms": round(tot_ms / n, 4),
                             "tflops": round(ach, 1), "frac": round(ach / peak, 3),
                             "share_of_step": round(tot_ms / prof_steps / ms_res, 4)})
        tot_ms, (kind, M, N, K), n = rows[0]
        ach = 2.0 * M * N * K / (tot_ms / n * 1e-3) / 1e12
        tr = traffic.get(f"{kind}|{M}|{N}|{K}")
        roof = {"bound": "tensor", "achieved": round(ach, 2), "peak": peak, "unit": "TFLOP/s",
                "frac": round(ach / peak, 4), "traffic": (tr or {}).get("dram_bytes"),
                "kernel": f"{kind} M={M} N={N} K={K}" + (f" [{tr['kernel']}]" if tr else ""), "launches_timed": n,
                "avg_launch_ms": round(tot_ms / n, 4), "share_of_step": round(tot_ms / prof_steps / ms_res, 4),
                "algorithmic_flops_per_launch": 2.0 * M * N * K,
                "timed_over": f"{prof_steps} meta-steps" + (" launched kernel by kernel (one lane) after the graph-replayed timed region" if graphs else ""),
                "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (of measured)" if peaks else "fallback 1.4 PF sustained (of fallback)",
                "traffic_source": "profiles/r1_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full" if tr else None}

    h2d = sum(sum(b[1][0].numel() * 4 for b in tr) + te[1][0].numel() * 4 for tr, te in host_tasks)
    h2d += sum((len(tr) + 1) * (INNER_B * 8 + 2 * INNER_B * (L_TGT + 1) * 8) for tr, te in host_tasks)
    out = {
        "metric": "meta-train frames/sec", "value": round(FRAMES_PER_STEP / (ms_res * 1e-3), 1), "unit": "frames/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_res, 3),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": WORKLOAD, "frames_per_step": FRAMES_PER_STEP, "parallelism": f"task-dp{world}",
                   "accents_per_rank": len(mine), "task_lanes": min(args.lanes, len(mine)), "gemm_path": gemm,
                   "cuda_graphs": bool(graphs),
                   "l2": "activations streamed per batch (several GB) >> 126 MB L2; no explicit flush"},
        "e2e": {"value": round(FRAMES_PER_STEP / (ms_e2e * 1e-3), 1), "unit": "frames/s",
                "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(len(mine) * 4 * 8),
                "ms_per_step": round(ms_e2e, 3)},
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "roofline_top_kernels": roof_all,
        "last_inner_test_loss": [round(i["loss"], 4) for i in loss_info][:2],
    }
    if rank == 0:
        out["ctc"] = ctc_bandwidth(be, peaks)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline_sample()
    if rank == 0:
        print(json.dumps(out), flush=True)
    if D.is_dist():
        torch.distributed.destroy_process_group()


def ctc_bandwidth(be, peaks):
    """Kernel 1 (CTC alpha-beta forward-backward, src/blstm_trainer.py:22,55-70): achieved HBM GB/s on the
    algorithmic bytes B*T'*C*(4+4) at the BASELINE shape (latency regime, 32 utterances on 148 SMs) and in the
    bandwidth regime (2048 and 8192 utterances: 0.77 / 3.1 GB per launch, far beyond L2), CUDA events over 10 launches each."""
    dev = be.device
    res = {}
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    for (T, B, C, L) in [(128, 32, 367, 34), (128, 2048, 367, 34), (128, 8192, 367, 34)]:
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
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            run()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / 10
        gbs = T * B * C * 8 / us / 1e3
        res[f"B{B}_T{T}_C{C}_L{L}"] = {"us": round(us, 1), "GB/s": round(gbs, 1), "frac_of_hbm_peak": round(gbs / hbm, 3)}
    tj = ROOT / "profiles" / "r1_traffic.json"
    if tj.exists():                  # DRAM bytes per launch of the same kernel from the committed ncu --set full capture
        for key, v in json.loads(tj.read_text()).get("ctc", {}).items():
            if key in res and isinstance(v, dict):
                res[key]["traffic"] = v.get("dram_bytes")
                res[key]["algorithmic_bytes"] = v.get("algorithmic_bytes")
    return res


# ================================================================================================= CPU baseline / reference arm
def _port_setup(seed=531):
    from oracle import port
    cfg = port.NetCfg(dropout=0.1, pos_dropout=0.1)
    sd = port.init_state_dict(cfg, seed=seed)
    ml = port.MetaLearner(sd, cfg, algo="fomaml", k=1.0, warmup=25000, eps_ls=0.2, training=True)
    return port, cfg, ml


def cpu_baseline_sample(B=8, steps=2):
    """Oracle port (kind 'port') on the host cores: one accent's inner-train + inner-test batch + meta-update
    at a reduced inner batch, frames/s."""
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    port, cfg, ml = _port_setup()
    gen = torch.Generator().manual_seed(1)
    mk = lambda: synth_batch(gen, B=B)
    ml.meta_step([([mk()], mk())])                          # warm-up
    t0 = time.time()
    for _ in range(steps):
        ml.meta_step([([mk()], mk())])
    dt = (time.time() - t0) / steps
    return {"value": round(2 * B * T_FRAMES / dt, 1), "unit": "frames/s", "cores": cores, "kind": "port",
            "sample": f"1 accent x (1 inner-train + 1 inner-test batch of {B} x T{T_FRAMES}) + meta-update per step, "
                      f"{steps} timed steps after 1 warm-up; torch {torch.__version__} CPU kernels"}


def run_reference(args):
    """--impl reference: the reference algorithm's CPU implementation (oracle port) on all host cores; each step
    is a bounded sample of the workload (1 of the 8 accents at inner batch 4)."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    B = 4
    port, cfg, ml = _port_setup()
    gen = torch.Generator().manual_seed(1)
    mk = lambda: synth_batch(gen, B=B)
    for _ in range(args.warmup):
        ml.meta_step([([mk()], mk())])
    t0 = time.time()
    for _ in range(args.steps):
        ml.meta_step([([mk()], mk())])
    dt = (time.time() - t0) / max(args.steps, 1)
    frames = 2 * B * T_FRAMES
    v = round(frames / dt, 1)
    sample = (f"per step: 1 of the 8 accents, inner-train + inner-test batch of {B} x T{T_FRAMES} x 83 (fwd+bwd, clip, "
              f"nesterov SGD, accumulate) + noam-Adam meta-update over 24.9 M parameters")
    out = {"impl": "reference", "metric": "meta-train frames/sec", "value": v, "unit": "frames/s",
           "n_gpus": int(os.environ.get("WORLD_SIZE", args.gpus)), "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": round(dt * 1e3, 2), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
           "dtype": "f32", "data": "synthetic",
           "config": {"workload": WORKLOAD, "frames_per_step": frames, "sampled": True},
           "cpu_baseline": {"value": v, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
           "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", dest="no_cpu_baseline", action="store_true")
    ap.add_argument("--no-graphs", dest="no_graphs", action="store_true", help="launch every kernel from the host")
    ap.add_argument("--lanes", type=int, default=3,
                    help="accents of a rank's share that run concurrently on one GPU (asr_model.task_lanes)")
    ap.add_argument("--profile", default=None, help="write a per-entry-point CUDA-event time table to this file")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
