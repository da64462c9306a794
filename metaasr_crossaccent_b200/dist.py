"""Task-level data parallelism for the meta-step: one process per GPU, accents of a
meta-batch partitioned across ranks, ONE all-reduce of the flat update arena per meta-step.

The reference has no distributed code at all; this is the only collective on the path.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def env_world():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))


def init_from_env(backend=None):
    """Initialise torch.distributed from the torchrun environment (no-op for world size 1)."""
    rank, world, local = env_world()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kw["device_id"] = torch.device("cuda", local)     # eager communicator; symmetric-memory rendezvous needs the device
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, world, local


def is_dist():
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def rank():
    return dist.get_rank() if is_dist() else 0


def world_size():
    return dist.get_world_size() if is_dist() else 1


def partition_tasks(task_ids, meta_batch_size, r=None, w=None):
    """Every rank draws the same shuffle (same python `random` seed, pretrain.py:62) and takes
    task_ids[:meta_batch_size][rank::world]."""
    r = rank() if r is None else r
    w = world_size() if w is None else w
    return list(task_ids[:meta_batch_size])[r::w]


def all_reduce_sum_(flat: torch.Tensor):
    """SUM all-reduce of the flat update arena (its last slot carries the task counter)."""
    if is_dist():
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    return flat


def barrier():
    if is_dist():
        dist.barrier()
