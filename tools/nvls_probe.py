#!/usr/bin/env python3
"""All-reduce of the 100 MB update arena: NCCL vs the NVSwitch multicast path (symmetric memory, multimem.ld_reduce /
multimem.st).  torchrun --nproc-per-node N tools/nvls_probe.py"""
import os
import sys
import time
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    n = 24_900_000 // (4 * world) * (4 * world)
    import torch.distributed._symmetric_memory as symm_mem
    t = symm_mem.empty(n, dtype=torch.float32, device=dev)
    hdl = symm_mem.rendezvous(t, dist.group.WORLD)
    if rank == 0:
        print("multicast_ptr", hex(hdl.multicast_ptr), "has_multicast", getattr(hdl, "has_multicast_support", None),
              "world", hdl.world_size, flush=True)
    g = torch.Generator(device=dev).manual_seed(rank)
    src = torch.randn(n, device=dev, generator=g)
    ref = src.clone()
    dist.all_reduce(ref)
    gname = dist.group.WORLD.group_name

    def timeit(fn, reps=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    buf = src.clone()
    ms_nccl = timeit(lambda: dist.all_reduce(buf))
    t.copy_(src)
    torch.ops.symm_mem.multimem_all_reduce_(t, "sum", gname)
    err = float((t - ref).abs().max())
    ms_mm = timeit(lambda: torch.ops.symm_mem.multimem_all_reduce_(t, "sum", gname))
    # our kernel on the same multicast mapping
    from metaasr_crossaccent_b200 import _lib
    lib = _lib.init(local)
    per = n // world

    def ours():
        hdl.barrier(channel=0)
        rc = lib.masr_nvls_allreduce_f32(hdl.multicast_ptr, rank * per, (rank + 1) * per, torch.cuda.current_stream().cuda_stream)
        assert rc == 0
        hdl.barrier(channel=1)
    t.copy_(src)
    ours()
    err2 = float((t - ref).abs().max())
    ms_ours = timeit(ours)
    if rank == 0:
        print(f"N={world} {n * 4 / 1e6:.1f} MB: NCCL {ms_nccl:.3f} ms, torch multimem op {ms_mm:.3f} ms (max err {err:.2e}), "
              f"masr_nvls_allreduce_f32 {ms_ours:.3f} ms (max err {err2:.2e})", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
