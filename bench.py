#!/usr/bin/env python3
"""bench.py -- meta-train frames/s of the hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--dtype bf16|fp32]
                    [--config fomaml|reptile|multi|blstm_ctc] [--no-extras]

Default (`--config fomaml` = BASELINE configs[1], the headline line the driver records): a "step" is ONE FOMAML
meta-step of the fometa-hkust network (d512/h8/ff2048/2e4d, C=367, label smoothing 0.2, dropout 0.1): 8 synthetic
accents, meta_k = 1, inner batch 32 x 512 frames x 83-dim fbank+pitch, targets of 32 unigram150 ids -> per accent one
inner-train batch (fwd+bwd, clip, nesterov-SGD) and one inner-test batch (fwd+bwd, clip, accumulate), then all-reduce +
noam-Adam meta-update = 262 144 input frames per step.  Accents are partitioned over the ranks (strong scaling).

  value    : frames/s with the step's batches already resident in HBM (CUDA events, max over ranks)
  e2e      : same metric through the public drop-in API (get_trainer / run_task / run_batch) from pinned HOST buffers:
             host->device copies of every batch and the device->host read of the losses are inside the timed region
  roofline : the kernel shape with the largest TOTAL time in the step (no duration filter), algorithmic FLOPs / its
             average CUDA-event launch time, against the burst bf16 peak of MEASURED_PEAKS.json when the clock record
             shows no power cap (else the sustained one); `roofline_top_kernels` lists the next shapes
  phases   : CUDA-event split of a meta-step (inner-train / inner-test / all-reduce / Adam), one lane
  replica_check (N > 1): spread of the meta-weight checksum over the ranks (must be 0) and rel-L2 of the N-rank
             all-reduced meta-gradient against a sequential 1-rank replay of the same 8 accents
  other_configs : short runs of BASELINE configs 3 (Reptile, meta_k 4), 4 (multi-task joint CTC/attention, gradient DP),
             5 (VGG-BLSTM CTC-only) and the fp32 mode of the headline config; each is also reachable as --config
  ctc      : kernel 1 over the SURVEY 8(d) sweep, achieved HBM GB/s on the algorithmic bytes B*T'*C*8
  --impl reference : the reference's algorithm on the host CPU cores (oracle/port.py, the checker that is pinned to the
             live reference by tests/golden), one bounded sample per step: ONE of the 8 accents at the full inner batch.
"""
from __future__ import annotations

import argparse
import gc
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

N_ACCENTS, INNER_B, T_FRAMES, L_TGT, IDIM = 8, 32, 512, 32, 83
NET = "fometa-hkust transformer (d512 h8 ff2048 2enc 4dec, C=367)"
SHAPE = "inner batch 32 x T512 x 83-dim fbank, L=32 unigram150 ids"
WORKLOADS = {
    "fomaml": f"FOMAML meta-step, {NET}, 8 synthetic accents, meta_k 1, {SHAPE}",
    "reptile": f"Reptile meta-step, {NET}, 8 synthetic accents, meta_k 4 (+1 forward-only logging batch per accent), {SHAPE}",
    "multi": f"multi-task step (MultiASRInterface), joint CTC/attention ctc_weight 0.3, {NET}, one batch per GPU per step "
             f"(gradient all-reduce), {SHAPE}",
    "blstm_ctc": "VGG-BLSTM (3 x BLSTM-360 + projection, stock torch/cuDNN) CTC-only step with B200CTCLoss, "
                 "batch 32 x T512 x 83 -> T'=128, targets L+2=34, SGD nesterov",
}
ACCENTS = ["af", "au", "ca", "en", "in", "ir", "nz", "us"]


def hkust_config(dtype, gemm, dropout=0.1, graphs=True, lanes=1, meta=True, ctc_weight=0.0, nvls=True):
    am = {"idim": IDIM, "nheads": 8, "d_model": 512, "d_inner": 2048, "dropout": dropout, "tgt_share_weight": 1,
          "encoder": {"nlayers": 2}, "decoder": {"nlayers": 4}, "pos_dropout": dropout, "dtype": dtype, "gemm": gemm,
          "cuda_graphs": graphs, "task_lanes": lanes, "ctc_weight": ctc_weight, "nvls_meta_update": nvls,
          "strict_tcgen05": gemm == "umma"}        # the timed path must never drop to a CUDA-core kernel
    if meta:
        am.update({"inner_optimizer_cls": "SGD", "inner_optimizer_opt": {"momentum": 0.9, "nesterov": True},
                   "meta_opt_cls": "noam", "meta": {"optimizer_opt": {"k": 1.0, "warmup_steps": 25000}}})
    else:
        am.update({"optimizer_cls": "noam", "optimizer_opt": {"k": 1.0, "warmup_steps": 25000}})
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


def load_peaks():
    pk = ROOT / "MEASURED_PEAKS.json"
    return json.loads(pk.read_text()) if pk.exists() else {}


def timed(fn, steps, warmup, dev, world, be=None, profile=False):
    """W untimed steps, barrier + sync, K steps between two CUDA events on the launching stream, sync + barrier; MAX over
    ranks.  Returns (ms per step, our kernel launches per step, per-launch event table or None, host (t0, t1))."""
    from metaasr_crossaccent_b200 import dist as D
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    if world > 1:           # (world = 1 also for legs that only rank 0 runs inside a multi-rank job: no collective there)
        D.barrier()
    if be is not None:
        be.launches = 0
        be.prof = {} if profile else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        D.barrier()
    t1 = time.time()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
    prof = None
    launches = 0
    if be is not None:
        prof, be.prof = be.prof, None
        launches = be.launches // max(steps, 1)
    return float(ms) / max(steps, 1), launches, prof, (t0, t1)


# ================================================================================================= FOMAML / Reptile
def make_meta_solver(algo, meta_k, dtype, graphs, lanes, nvls=True):
    from metaasr_crossaccent_b200 import interfaces as I
    from metaasr_crossaccent_b200.trainer import get_trainer
    import random
    gemm = "umma" if dtype == "bf16" else "simt"
    id2accent = {a: a for a in ACCENTS + ["hk"]}
    paras = argparse.Namespace(pretrain_accents=ACCENTS, num_pretrain=N_ACCENTS, tgt_accent="hk", runs=0, seed=531,
                               meta_k=meta_k, meta_batch_size=N_ACCENTS, max_step=0, resume=False, algo=algo,
                               pretrain_suffix="bench", log_root=None)
    random.seed(531); torch.manual_seed(531)
    solver = get_trainer(I.FOMetaASRInterface, hkust_config(dtype, gemm, graphs=graphs, lanes=lanes, nvls=nvls), paras, id2accent)
    solver.set_model()
    return solver, gemm


def host_tasks_of(rank, world, meta_k, pin=True):
    from metaasr_crossaccent_b200 import dist as D
    mine = D.partition_tasks(list(range(N_ACCENTS)), N_ACCENTS, rank, world)
    gen = torch.Generator().manual_seed(531 + rank)
    return mine, [([(a, synth_batch(gen, pin=pin)) for _ in range(meta_k)], (a, synth_batch(gen, pin=pin))) for a in mine]


def clone_host(task):
    tr, te = task
    cl = lambda b: (b[0], (b[1][0], b[1][1].clone(), b[1][2], b[1][3].clone()))   # olens is mutated in place
    return [cl(b) for b in tr], cl(te)


def flops_of(key):
    kind = key[0]
    if kind == "attn_fwd":
        return 4.0 * key[1] * key[2] * key[3] * 64
    if kind == "attn_bwd":
        return 10.0 * key[1] * key[2] * key[3] * 64
    _, M, N, K = key
    return 2.0 * M * N * K


def key_name(key):
    if key[0].startswith("attn"):
        return f"{key[0]} BH={key[1]} Lq={key[2]} Lk={key[3]} d=64"
    return f"{key[0]} M={key[1]} N={key[2]} K={key[3]}"


def bench_meta(args, algo, meta_k, dtype, steps, warmup, detail, rank, world, dev):
    """One line for a FOMAML / Reptile meta-step configuration."""
    from metaasr_crossaccent_b200 import dist as D
    graphs = not args.no_graphs
    # the headline run takes the NVSwitch-multicast meta-update when there are several ranks; the short runs of the other
    # configurations (several solvers in one process) stay on the NCCL all-reduce
    solver, gemm = make_meta_solver(algo, meta_k, dtype, graphs, args.lanes, nvls=detail)
    eng, be = solver.asr_model.engine, solver.backend
    mine, host_tasks = host_tasks_of(rank, world, meta_k)
    frames = N_ACCENTS * (meta_k + 1) * INNER_B * T_FRAMES

    def prepared(task):
        tr, te = clone_host(task)
        mk = lambda b: (b[0], (eng.to_device(eng.prepare_batch(*b[1])), None, [None] * INNER_B, None))
        return [mk(b) for b in tr], mk(te)

    dev_tasks = [prepared(t) for t in host_tasks]            # inputs resident in HBM
    torch.cuda.synchronize()
    step_resident = lambda: solver.meta_step_on_tasks(dev_tasks, global_task_count=N_ACCENTS)

    last_info = {}

    nxt = {}

    def step_e2e():
        # like a prefetching loader: the NEXT step's host batches are handed over with the current ones, so their
        # host->device copies (copy stream) run under this step's compute; every step still copies all of its inputs
        cur = nxt.pop("t", None) or [clone_host(t) for t in host_tasks]
        nxt["t"] = [clone_host(t) for t in host_tasks]
        solver.meta_step_on_tasks(cur, global_task_count=N_ACCENTS, next_tasks=nxt["t"])
        last_info["i"] = solver.flush_train_info()           # device->host read of the step's losses
        return last_info["i"]

    sampler = ClockSampler(dev.index)
    sampler.start()
    ms_res, launches, prof, (t0, t1) = timed(step_resident, steps, warmup, dev, world, be, profile=not graphs)
    clocks = sampler.stop(t0, t1)
    ms_e2e, _, _, _ = timed(step_e2e, steps, max(1, warmup // 2), dev, world, be)
    solver.flush_train_info()
    h2d = sum(sum(b[1][0].numel() * 4 for b in tr) + te[1][0].numel() * 4 for tr, te in host_tasks)
    h2d += sum((len(tr) + 1) * (INNER_B * 8 + 2 * INNER_B * (L_TGT + 1) * 8) for tr, te in host_tasks)
    out = {
        "metric": "meta-train frames/sec", "value": round(frames / (ms_res * 1e-3), 1), "unit": "frames/s",
        "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": round(ms_res, 3),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": dtype, "data": "synthetic",
        "config": {"workload": WORKLOADS[algo], "frames_per_step": frames, "parallelism": f"task-dp{world}",
                   "accents_per_rank": len(mine), "task_lanes": min(args.lanes, len(mine)), "gemm_path": gemm,
                   "cuda_graphs": bool(graphs),
                   "l2": "activations streamed per batch (several GB) >> 126 MB L2; no explicit flush"},
        "e2e": {"value": round(frames / (ms_e2e * 1e-3), 1), "unit": "frames/s",
                "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(len(mine) * 8 * 8),
                "ms_per_step": round(ms_e2e, 3)},
        "gpu_launches": int(launches), "clocks": clocks,
        "last_inner_test_loss": [round(i["loss"], 4) for i in last_info.get("i", [])][:2],
    }
    if world > 1:
        out["config"]["meta_update"] = ("one kernel over NVSwitch multicast memory (masr_nvls_reduce_adam: sum in the switch, "
                                        "sharded Adam, new weights multicast)" if getattr(solver, "_nvls", None) is not None
                                        else "NCCL all-reduce + replicated Adam")
    if not detail:
        if graphs:                                # graph replays bypass the host-side launch counter: count one eager step
            eng.use_graphs = False                # (one lane: the other lanes' engines keep their graphs)
            solver.config["asr_model"]["task_lanes"] = 1
            _, out["gpu_launches"], _, _ = timed(step_resident, 1, 0, dev, world, be)
        del solver, dev_tasks
        return out

    # ---- per-launch CUDA-event times of the tensor-core kernels (GEMM / conv / attention shapes), launched kernel by
    # kernel on ONE lane without the side stream, behind a device-side spin so that the launch queue is always full:
    # the event deltas are then GPU time, not host launch gaps (small launches take 4-15 us)
    lanes_cfg = solver.config["asr_model"].get("task_lanes", 1)
    solver.config["asr_model"]["task_lanes"] = 1
    g_cfg, ms_cfg = eng.use_graphs, eng.multi_stream
    eng.use_graphs, eng.multi_stream = False, False
    fb = eng.forward_backward

    def fb_behind_spin(db):
        torch.cuda._sleep(int(8e6))               # ~4 ms at 1.9 GHz: the host runs ahead of the GPU for the whole batch
        return fb(db)
    eng.forward_backward = fb_behind_spin
    step_resident()
    prof_steps = 2
    _, launches_eager, prof, _ = timed(step_resident, prof_steps, 0, dev, world, be, profile=True)
    eng.forward_backward = fb
    # what an event-bracketed launch costs when the kernel itself does nothing (one thread, no work): the two event records
    # and the launch gap that PDL / graph replay hide in the real step.  Small launches (decoder GEMMs: ~4 us inside the
    # replayed graph, tools/chain_probe.py) would otherwise be charged ~2.5x their cost in the step when kernels are
    # ranked by total time.
    torch.cuda._sleep(int(4e6))
    be.prof = {}
    for _ in range(256):
        be._timed_call(("null",), "masr_seed_bump", be._seed_t.data_ptr(), 0, be.stream)
    torch.cuda.synchronize()
    null_evs = be.prof.pop(("null",))
    be.prof = None
    null_ms = statistics.median(a.elapsed_time(b) for a, b in null_evs)
    if not graphs:
        launches_eager = launches
    out["gpu_launches"] = int(launches_eager)
    out["gpu_launches_note"] = ("kernel launches of one meta-step counted on an eager one-lane pass; the timed multi-lane schedule "
                                "issues the weight-gradient GEMMs of each lane as grouped launches (about 11 % fewer)")
    # ---- phase split of a meta-step on one lane (graph replay as configured): CUDA events at the phase boundaries
    eng.use_graphs, eng.multi_stream = g_cfg, ms_cfg
    step_resident()
    solver._phase_log = []
    for _ in range(3):
        step_resident()
    torch.cuda.synchronize()
    log, solver._phase_log = solver._phase_log, None
    ph = {"inner_train": 0.0, "inner_test": 0.0, "all_reduce": 0.0, "meta_adam": 0.0}
    for (ta, ea), (tb, eb) in zip(log[:-1], log[1:]):
        dt = ea.elapsed_time(eb) / 3
        if tb == "train1":
            ph["inner_train"] += dt
        elif tb == "test1":
            ph["inner_test"] += dt
        elif tb == "reduce1":
            ph["all_reduce"] += dt
        elif tb == "adam1":
            ph["meta_adam"] += dt
    out["phases_ms_one_lane"] = {k: round(v, 3) for k, v in ph.items()}
    solver.config["asr_model"]["task_lanes"] = lanes_cfg

    peaks = load_peaks()
    capped = "sw_power_cap" in (clocks.get("reasons") or [])
    peak = float(peaks.get("bf16_tflops_sustained" if capped else "bf16_tflops", 1400.0 if capped else 1650.0))
    peak_src = ("MEASURED_PEAKS.json " if peaks else "fallback ") + \
               ("bf16_tflops_sustained (sw_power_cap seen during the timed region)" if capped else
                "bf16_tflops (burst: no cap in the clock record, kernels timed one by one)")
    if prof:
        traffic = {}
        tj = ROOT / "profiles" / "r1_traffic.json"
        if tj.exists():
            traffic = json.loads(tj.read_text()).get("per_launch", {})
        # ranked by time NET of the null-launch cost (floor: a quarter of the raw time); achieved TFLOP/s and frac are
        # computed from the RAW event time of the launch (conservative)
        net = lambda evs: sum(max(a.elapsed_time(b) - null_ms, 0.25 * a.elapsed_time(b)) for a, b in evs)
        rows = sorted(((net(evs), sum(a.elapsed_time(b) for a, b in evs), key, len(evs)) for key, evs in prof.items()),
                      reverse=True)
        tot_all = sum(r[0] for r in rows)
        top = []
        for net_ms, tot_ms, key, n in rows[:10]:
            ach = flops_of(key) / (tot_ms / n * 1e-3) / 1e12
            top.append({"kernel": key_name(key), "launches": n, "avg_launch_ms": round(tot_ms / n, 4),
                        "avg_launch_ms_net": round(net_ms / n, 4), "tflops": round(ach, 1), "frac": round(ach / peak, 3),
                        "share_of_tensor_time": round(net_ms / tot_all, 4)})
        net_ms, tot_ms, key, n = rows[0]
        ach = flops_of(key) / (tot_ms / n * 1e-3) / 1e12
        tr = traffic.get("|".join(str(v) for v in key))
        total_flops = sum(flops_of(k) * len(v) for k, v in prof.items()) / prof_steps
        out["roofline"] = {
            "bound": "tensor", "achieved": round(ach, 2), "peak": peak, "unit": "TFLOP/s", "frac": round(ach / peak, 4),
            "traffic": (tr or {}).get("dram_bytes"),
            "kernel": key_name(key) + (f" [{tr['kernel']}]" if tr else ""), "launches_timed": n,
            "avg_launch_ms": round(tot_ms / n, 4), "share_of_tensor_time": round(net_ms / tot_all, 4),
            "algorithmic_flops_per_launch": flops_of(key), "null_launch_ms": round(null_ms, 4),
            "selected_by": "largest total CUDA-event time over all tensor-core launch shapes of the step, net of the "
                           "event-bracketed null-launch time (no duration filter); achieved / frac use the raw time",
            "timed_over": f"{prof_steps} meta-steps launched kernel by kernel behind a device-side spin (one lane, no side stream)",
            "peak_source": peak_src,
            "traffic_source": "profiles/r1_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full" if tr else None}
        out["roofline_top_kernels"] = top
        out["step_tensor"] = {"flops_per_step": total_flops, "tflops": round(total_flops / (ms_res * 1e-3) / 1e12, 1),
                              "frac_of_peak": round(total_flops / (ms_res * 1e-3) / 1e12 / peak, 4),
                              "note": "this rank's tensor-core FLOPs per meta-step / ms_per_step (whole step incl. all-reduce and Adam)"}

    # ---- replicas: bit-identical meta weights over the ranks; N-rank meta-gradient vs a sequential 1-rank replay
    if world > 1:
        out["replica_check"] = replica_check(solver, eng, rank, world, dev, meta_k)
    del solver, dev_tasks
    return out


def replica_check(solver, eng, rank, world, dev, meta_k):
    from metaasr_crossaccent_b200 import dist as D
    cs = solver.replica_checksum()
    lo, hi = cs.clone(), cs.clone()
    torch.distributed.all_reduce(lo, op=torch.distributed.ReduceOp.MIN)
    torch.distributed.all_reduce(hi, op=torch.distributed.ReduceOp.MAX)
    spread = float((hi - lo).abs().max())
    # meta-gradient of the distributed step (dropout off, kernel-by-kernel: dropout streams depend on the launch order)
    cfg = eng.cfg
    pd, ppd, g_cfg = cfg.dropout, cfg.pos_dropout, eng.use_graphs
    lanes_cfg = solver.config["asr_model"].get("task_lanes", 1)
    cfg.dropout = cfg.pos_dropout = 0.0
    eng.use_graphs = False
    solver.config["asr_model"]["task_lanes"] = 1
    n = eng.layout.total

    def accumulate(tasks):
        solver._upd_flat.zero_()
        solver._counter = 0
        for tr, te in tasks:
            solver.run_task(tr)
            solver.inner_test(te)
        solver._ring_sizes = []

    _, mine_tasks = host_tasks_of(rank, world, meta_k, pin=False)
    accumulate([clone_host(t) for t in mine_tasks])
    u_own = solver._upd_flat[:n].clone()                 # this rank's share before the collective
    solver._reduce_updates()
    u_dist = solver._upd_flat[:n].clone()
    every, owner = [], []
    for r in range(world):
        ts = host_tasks_of(r, world, meta_k, pin=False)[1]
        every += ts
        owner += [r] * len(ts)
    # sequential replay on THIS rank, keeping the share of every rank apart: which rank's share differs, and where
    shares = [torch.zeros(n, device=dev) for _ in range(world)] if world <= 8 else None
    solver._upd_flat.zero_()
    solver._counter = 0
    prev = torch.zeros(n, device=dev)
    for t, r in zip(every, owner):
        tr, te = clone_host(t)
        solver.run_task(tr)
        solver.inner_test(te)
        cur = solver._upd_flat[:n].clone()
        if shares is not None:
            shares[r] += cur - prev
        prev = cur
    solver._ring_sizes = []
    u_seq = solver._upd_flat[:n].clone()
    solver._upd_flat.zero_()
    solver._counter = 0
    rel = float((u_dist - u_seq).norm() / u_seq.norm().clamp_min(1e-30))
    per_rank, worst_tensor = None, None
    if shares is not None:
        mine_seq = shares[rank]
        own_rel = torch.tensor([float((u_own - mine_seq).norm() / mine_seq.norm().clamp_min(1e-30))], device=dev)
        allr = [torch.zeros_like(own_rel) for _ in range(world)]
        torch.distributed.all_gather(allr, own_rel)
        per_rank = [round(float(v), 10) for v in allr]
        worst = max(((float((eng.layout.view(u_own, nm) - eng.layout.view(mine_seq, nm)).norm() /
                             eng.layout.view(mine_seq, nm).norm().clamp_min(1e-30)), nm) for nm in eng.layout.offsets))
        wl = [None] * world
        torch.distributed.all_gather_object(wl, (round(worst[0], 8), worst[1]))
        worst_tensor = max(wl)
    cfg.dropout, cfg.pos_dropout, eng.use_graphs = pd, ppd, g_cfg
    solver.config["asr_model"]["task_lanes"] = lanes_cfg
    torch.cuda.synchronize(); D.barrier()
    return {"meta_weight_checksum_spread_over_ranks": spread,
            "meta_grad_rel_l2_vs_sequential_replay": rel,
            "own_share_rel_l2_vs_its_replay_on_this_rank_by_rank": per_rank, "worst_tensor_of_any_rank": worst_tensor,
            "note": "checksum = (sum, sum of squares) of _original_flat in float64 after the timed steps, max - min over ranks; "
                    "replay = all 8 accents run on this rank alone (dropout 0), bf16 split-K reductions use fp32 atomics "
                    "(run-to-run order noise ~1e-7 when the reduction order repeats; when it does not, the last-bit "
                    "difference of the inner-train step flips bf16 roundings of the inner-test forward and the two bf16 "
                    "realisations of an accent's share differ at the bf16 level, 1e-3..1e-2 -- tools/flaky_probe.py shows the "
                    "chain element by element); own_share...: every rank's own share against the same accents inside its "
                    "sequential replay on the same GPU"}


# ================================================================================================= multi-task (config 4)
def bench_multi(args, dtype, steps, warmup, rank, world, dev):
    from metaasr_crossaccent_b200 import interfaces as I
    from metaasr_crossaccent_b200.trainer import get_trainer
    import random
    gemm = "umma" if dtype == "bf16" else "simt"
    id2accent = {a: a for a in ACCENTS + ["hk"]}
    paras = argparse.Namespace(pretrain_accents=ACCENTS, num_pretrain=N_ACCENTS, tgt_accent="hk", runs=0, seed=531,
                               meta_k=None, meta_batch_size=None, max_step=0, resume=False, algo="multi",
                               pretrain_suffix="bench", log_root=None)
    random.seed(531); torch.manual_seed(531)
    solver = get_trainer(I.MultiASRInterface, hkust_config(dtype, gemm, graphs=not args.no_graphs, meta=False, ctc_weight=0.3),
                         paras, id2accent)
    solver.set_model()
    eng, be = solver.asr_model.engine, solver.backend
    gen = torch.Generator().manual_seed(977 + rank)
    host = [(rank % N_ACCENTS, synth_batch(gen, pin=True)) for _ in range(4)]      # this rank's own batches
    dev_items = [(a, (eng.to_device(eng.prepare_batch(b[0], b[1].clone(), b[2], b[3].clone())), None, [None] * INNER_B, None))
                 for a, b in host]
    torch.cuda.synchronize()
    it = {"i": 0}

    def step_resident():
        it["i"] += 1
        solver.multi_step(dev_items[it["i"] % len(dev_items)], sync=False)

    def step_e2e():
        it["i"] += 1
        a, b = host[it["i"] % len(host)]
        return solver.multi_step((a, (b[0], b[1].clone(), b[2], b[3].clone())), sync=True)     # D2H read of the loss

    sampler = ClockSampler(dev.index)
    sampler.start()
    ms_res, launches, _, (t0, t1) = timed(step_resident, steps, warmup, dev, world, be)
    clocks = sampler.stop(t0, t1)
    ms_e2e, _, _, _ = timed(step_e2e, steps, max(1, warmup // 2), dev, world, be)
    info = eng.read_stats()
    if eng.use_graphs:                            # graph replays bypass the host-side launch counter: count one eager step
        eng.use_graphs = False
        _, launches, _, _ = timed(step_resident, 1, 0, dev, world, be)
    frames = world * INNER_B * T_FRAMES
    out = {"metric": "meta-train frames/sec", "value": round(frames / (ms_res * 1e-3), 1), "unit": "frames/s",
           "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": round(ms_res, 3), "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": dtype, "data": "synthetic",
           "config": {"workload": WORKLOADS["multi"], "frames_per_step": frames, "parallelism": f"grad-dp{world}",
                      "gemm_path": gemm, "cuda_graphs": bool(eng.use_graphs), "ctc_weight": 0.3,
                      "l2": "activations streamed per batch >> 126 MB L2; no explicit flush"},
           "e2e": {"value": round(frames / (ms_e2e * 1e-3), 1), "unit": "frames/s",
                   "h2d_bytes_per_step": int(INNER_B * T_FRAMES * IDIM * 4 + INNER_B * 8 * (3 + 2 * (L_TGT + 1))),
                   "d2h_bytes_per_step": 64, "ms_per_step": round(ms_e2e, 3)},
           "gpu_launches": int(launches), "clocks": clocks,
           "last_loss": {k: round(v, 4) for k, v in info.items()}}
    del solver, dev_items
    return out


# ================================================================================================= VGG-BLSTM CTC (config 5)
class VGGBLSTM(torch.nn.Module):
    """Stock-torch network of config/blstm/mono-test.yaml (src/modules/encoder.py VGGExtractor + RNNP, src/model/blstm/
    mono_blstm.py head): the CTC kernel is the only part of config 5 this build replaces."""

    def __init__(self, idim=83, enc_dim=360, proj_dim=360, odim=367, nlayers=3):
        super().__init__()
        nn = torch.nn
        self.vgg = nn.Sequential(nn.Conv2d(1, 64, 3, padding=1), nn.ReLU(), nn.Conv2d(64, 64, 3, padding=1), nn.ReLU(),
                                 nn.MaxPool2d(2, stride=2, ceil_mode=True),
                                 nn.Conv2d(64, 128, 3, padding=1), nn.ReLU(), nn.Conv2d(128, 128, 3, padding=1), nn.ReLU(),
                                 nn.MaxPool2d(2, stride=2, ceil_mode=True))
        d = 128 * ((((idim + 1) // 2) + 1) // 2)
        self.rnns = nn.ModuleList([nn.LSTM(d if i == 0 else proj_dim, enc_dim, batch_first=True, bidirectional=True)
                                   for i in range(nlayers)])
        self.bts = nn.ModuleList([nn.Linear(2 * enc_dim, proj_dim) for _ in range(nlayers)])
        self.head = nn.Linear(proj_dim, odim)

    def forward(self, x):
        h = self.vgg(x.unsqueeze(1)).transpose(1, 2).flatten(2)
        for i, (rnn, bt) in enumerate(zip(self.rnns, self.bts)):
            h = bt(rnn(h)[0])
            if i + 1 < len(self.rnns):
                h = torch.tanh(h)
        return self.head(h)


def bench_blstm_ctc(args, steps, warmup, rank, world, dev):
    from metaasr_crossaccent_b200.ctc import B200CTCLoss
    torch.manual_seed(531)
    model = VGGBLSTM().to(dev)
    opt = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.9, nesterov=True)
    ctc = B200CTCLoss(blank=0, reduction="mean", zero_infinity=True)
    gen = torch.Generator().manual_seed(531)
    x, ilens, ys, olens = synth_batch(gen, pin=True)
    eos = torch.tensor([366])
    y_true = torch.cat([torch.cat([eos, y, eos]) for y in ys])                 # blstm_trainer.py:56-61
    tl = olens + 2
    il = torch.ceil(torch.ceil(ilens.float() / 2) / 2).long()
    xd, yd = x.to(dev), y_true.to(dev)
    ev = {"ctc": []}

    def run(xin, ytrue):
        pred = torch.nn.functional.log_softmax(model(xin), dim=-1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        loss = ctc(pred.transpose(0, 1).contiguous(), ytrue, il, tl)
        e1.record()
        ev["ctc"].append((e0, e1))
        opt.zero_grad()
        loss.backward()
        opt.step()
        return loss

    step_res = lambda: run(xd, yd)
    step_e2e = lambda: float(run(x.to(dev, non_blocking=True), y_true.to(dev, non_blocking=True)))
    sampler = ClockSampler(dev.index)
    sampler.start()
    ms_res, _, _, (t0, t1) = timed(step_res, steps, warmup, dev, 1)
    clocks = sampler.stop(t0, t1)
    ctc_ms = statistics.median(a.elapsed_time(b) for a, b in ev["ctc"][-steps:])
    ms_e2e, _, _, _ = timed(step_e2e, steps, 1, dev, 1)
    frames = INNER_B * T_FRAMES
    return {"metric": "meta-train frames/sec", "value": round(frames / (ms_res * 1e-3), 1), "unit": "frames/s", "n_gpus": 1,
            "steps": steps, "warmup": warmup, "ms_per_step": round(ms_res, 3), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOADS["blstm_ctc"], "frames_per_step": frames,
                       "note": "network = stock torch (cuDNN LSTM / conv): library code; ours = the CTC loss only"},
            "e2e": {"value": round(frames / (ms_e2e * 1e-3), 1), "unit": "frames/s",
                    "h2d_bytes_per_step": int(x.numel() * 4 + y_true.numel() * 8), "d2h_bytes_per_step": 4,
                    "ms_per_step": round(ms_e2e, 3)},
            "gpu_launches": 2, "clocks": clocks,
            "ctc_loss_call_ms": round(ctc_ms, 4), "ctc_share_of_step": round(ctc_ms / ms_res, 4)}


# ================================================================================================= CTC sweep (kernel 1)
def ctc_bandwidth(be, peaks, full=True):
    """Kernel 1 (CTC alpha-beta forward-backward, src/blstm_trainer.py:22,55-70): achieved HBM GB/s on the algorithmic
    bytes B*T'*C*(4+4) over the SURVEY 8(d) sweep B x (T', L): (128, 34) is the BASELINE shape (T=512), (375, 102) and
    (750, 152) are the reference's longest training utterances (max_ilen 1500 / 3000).  CUDA events over 10 launches."""
    dev = be.device
    res = {}
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    shapes = [(128, 32, 34), (128, 2048, 34), (128, 8192, 34)]
    if full:
        shapes = [(T, B, L) for (T, L) in ((128, 34), (375, 102), (750, 152)) for B in (32, 128, 512, 2048)] + [(128, 8192, 34)]
    C = 367
    for (T, B, L) in shapes:
        if T * B * C * 8 > 12e9:
            continue
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
        del lg, grad, ws
    tj = ROOT / "profiles" / "r1_traffic.json"
    if tj.exists():                  # DRAM bytes per launch of the same kernel from the committed ncu --set full capture
        for key, v in json.loads(tj.read_text()).get("ctc", {}).items():
            if key in res and isinstance(v, dict):
                res[key]["traffic"] = v.get("dram_bytes")
                res[key]["algorithmic_bytes"] = v.get("algorithmic_bytes")
    return res


# ================================================================================================= ours: driver
def run_ours(args):
    from metaasr_crossaccent_b200 import dist as D
    assert torch.cuda.is_available(), "bench.py measures the CUDA path; there is no CPU fallback"
    rank, world, local = D.init_from_env("nccl")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)

    def free():
        gc.collect()
        torch.cuda.empty_cache()

    if args.config == "blstm_ctc":
        out = bench_blstm_ctc(args, args.steps, args.warmup, rank, world, dev) if rank == 0 else None
    elif args.config == "multi":
        out = bench_multi(args, args.dtype, args.steps, args.warmup, rank, world, dev)
    else:
        meta_k = 1 if args.config == "fomaml" else 4
        out = bench_meta(args, args.config, meta_k, args.dtype, args.steps, args.warmup, True, rank, world, dev)
        free()
        if args.config == "fomaml" and not args.no_extras:
            # short runs of the other BASELINE configurations (each also reachable as --config X)
            ks, kw = max(2, min(args.steps, 4)), 3
            extra = {}
            extra["reptile_meta_k4"] = bench_meta(args, "reptile", 4, args.dtype, ks, kw, False, rank, world, dev); free()
            extra["multi_joint_ctc_attention"] = bench_multi(args, args.dtype, max(ks, 8), kw, rank, world, dev); free()
            if args.dtype == "bf16":
                extra["fomaml_fp32_mode"] = bench_meta(args, "fomaml", 1, "fp32", 2, 3, False, rank, world, dev); free()
            if rank == 0:
                extra["blstm_ctc"] = bench_blstm_ctc(args, 8, kw, rank, world, dev); free()
            keep = ("value", "unit", "ms_per_step", "e2e", "scaling", "dtype", "gpu_launches", "steps", "warmup", "config",
                    "ctc_share_of_step", "ctc_loss_call_ms", "last_loss")
            out["other_configs"] = {k: {kk: v[kk] for kk in keep if kk in v} for k, v in extra.items()}
        if rank == 0:
            from metaasr_crossaccent_b200 import ops
            be = ops.CudaBackend(dev, torch.bfloat16, gemm="umma")
            out["ctc"] = ctc_bandwidth(be, load_peaks(), full=not args.no_extras)
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline_sample()
    if rank == 0 and out is not None:
        print(json.dumps(out), flush=True)
    if D.is_dist():
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


# ================================================================================================= CPU baseline / reference arm
def _port_setup(algo="fomaml", seed=531):
    from oracle import port
    cfg = port.NetCfg(dropout=0.1, pos_dropout=0.1)
    sd = port.init_state_dict(cfg, seed=seed)
    ml = port.MetaLearner(sd, cfg, algo=algo, k=1.0, warmup=25000, eps_ls=0.2, training=True)
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
    """--impl reference: the reference algorithm's CPU implementation (oracle port) on all host cores; each step is a
    bounded sample of the workload: ONE of the 8 accents at the FULL inner batch (32 x T512), so the sampled step differs
    from the timed workload only in the accent count (frames/s is per-frame work, independent of it)."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    B = INNER_B
    gen = torch.Generator().manual_seed(1)
    mk = lambda: synth_batch(gen, B=B)
    if args.config in ("fomaml", "reptile"):
        meta_k = 1 if args.config == "fomaml" else 4
        port, cfg, ml = _port_setup(args.config)
        step = lambda: ml.meta_step([([mk() for _ in range(meta_k)], mk())])
        frames = (meta_k + 1) * B * T_FRAMES
        sample = (f"per step: 1 of the 8 accents, {meta_k} inner-train + 1 inner-test batch of {B} x T{T_FRAMES} x 83 (fwd+bwd, "
                  f"clip, nesterov SGD, accumulate) + noam-Adam meta-update over 24.9 M parameters")
    elif args.config == "multi":
        port, cfg, ml = _port_setup("fomaml")
        step = lambda: ml.multi_step(mk())
        frames = B * T_FRAMES
        sample = f"per step: one multi-task step (attention objective) on a batch of {B} x T{T_FRAMES} x 83 + noam-Adam"
    else:
        model = VGGBLSTM()
        opt = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.9, nesterov=True)
        x, ilens, ys, olens = mk()
        eos = torch.tensor([366])
        y_true = torch.cat([torch.cat([eos, y, eos]) for y in ys])
        il = torch.ceil(torch.ceil(ilens.float() / 2) / 2).long()

        def step():
            pred = torch.nn.functional.log_softmax(model(x), dim=-1)
            loss = torch.nn.functional.ctc_loss(pred.transpose(0, 1), y_true, il, olens + 2, blank=0, reduction="mean",
                                                zero_infinity=True)
            opt.zero_grad(); loss.backward(); opt.step()
        frames = B * T_FRAMES
        sample = f"per step: VGG-BLSTM CTC step (stock torch CPU) on a batch of {B} x T{T_FRAMES} x 83"
    for _ in range(args.warmup):
        step()
    t0 = time.time()
    for _ in range(args.steps):
        step()
    dt = (time.time() - t0) / max(args.steps, 1)
    v = round(frames / dt, 1)
    out = {"impl": "reference", "metric": "meta-train frames/sec", "value": v, "unit": "frames/s",
           "n_gpus": int(os.environ.get("WORLD_SIZE", args.gpus)), "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": round(dt * 1e3, 2), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
           "dtype": "f32", "data": "synthetic",
           "config": {"workload": WORKLOADS[args.config], "frames_per_step": frames, "sampled": True},
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
    ap.add_argument("--config", default="fomaml", choices=list(WORKLOADS))
    ap.add_argument("--no-extras", dest="no_extras", action="store_true",
                    help="headline line only: skip the short runs of the other configs and the long CTC sweep")
    ap.add_argument("--no-cpu-baseline", dest="no_cpu_baseline", action="store_true")
    ap.add_argument("--no-graphs", dest="no_graphs", action="store_true", help="launch every kernel from the host")
    ap.add_argument("--lanes", type=int, default=4,
                    help="accents of a rank's share that run concurrently on one GPU (asr_model.task_lanes)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
