"""Input pipeline of the hot path: drop-in for `src/io/dataset.py` -- same on-disk format, same
batches, different plumbing.

On-disk format (unchanged, Appendix B): `<dir>/feat.dat` (npy format, opened as a memmap) or `feat.npy`
[sum(ilens), 83] fp32, `ilens.npy`, `label.npy` [sum(olens)] int, `olens.npy`.

Reference behaviour kept bit for bit (`tests/test_data_host.py` pins it against the live reference's loader):

  * `BucketSampler` (dataset.py:35-119): 1-frame buckets (`np.digitize(ilens, bins, right=True)`), the bucket list
    shuffled once with `random.shuffle`, every bucket shuffled with `np.random.shuffle` when the iteration reaches
    it, half batch size beyond `half_batch_ilen`;
  * `collate_fn` (dataset.py:21-33): stable sort by ilen (descending), zero padding to the longest utterance,
    `ys` as a list of 1-D label tensors;
  * `get_loader` / `DataContainer.get_item` (dataset.py:158-277): arguments, iteration order, the re-created loader
    on `StopIteration`, `np.random.randint` accent draw for the multi-task interface.

What changes: no `torch.utils.data.DataLoader` worker processes and no per-utterance tensors / `pad_sequence`.
A batch is assembled by ONE pass of row-range copies from the memmap straight into a PINNED buffer (the
reference pins after collating: one extra copy), on a background thread that runs `prefetch`
batches ahead of the consumer; with `device=` the loader also issues the host->device copy on its own CUDA stream,
so the trainer's `run_batch` finds the features resident (engine.to_device accepts both).  Index sampling stays on
the CONSUMER side in the reference's order, so `prefetch=0` consumes the numpy / python RNG streams exactly like
the reference with `num_workers=0`; a prefetching loader draws its indices `prefetch` batches early (the reference
with worker processes does the same, two batches per worker).
"""
from __future__ import annotations

import queue
import random
import threading
from pathlib import Path

import numpy as np
import torch

BUCKET_SIZE = 1          # dataset.py:11
ILEN_MIN = 2             # dataset.py:12
ILEN_MAX = 10000         # dataset.py:13


class CommonVoiceDataset:
    """dataset.py:124-156: flat feature / label arrays with prefix-sum pointers."""

    def __init__(self, data_dir, is_memmap):
        data_dir = Path(data_dir)
        if is_memmap:
            self.feat = np.load(data_dir.joinpath('feat').with_suffix('.dat'), mmap_mode='r')
        else:
            self.feat = np.load(data_dir.joinpath('feat').with_suffix('.npy'))
        self.ilens = np.load(data_dir.joinpath('ilens.npy'))
        self.iptr = np.zeros(len(self.ilens) + 1, dtype=int)
        self.ilens.cumsum(out=self.iptr[1:])
        self.label = np.load(data_dir.joinpath('label.npy'))
        self.olens = np.load(data_dir.joinpath('olens.npy'))
        self.optr = np.zeros(len(self.olens) + 1, dtype=int)
        self.olens.cumsum(out=self.optr[1:])
        assert len(self.ilens) == len(self.olens), "Number of samples should be the same in features and labels"

    def __len__(self):
        return len(self.ilens)

    def __getitem__(self, idx):
        return {'feat': torch.as_tensor(np.asarray(self.feat[self.iptr[idx]:self.iptr[idx + 1], :])),
                'ilen': torch.as_tensor(self.ilens[idx]),
                'label': torch.as_tensor(self.label[self.optr[idx]:self.optr[idx + 1]]),
                'olen': torch.as_tensor(self.olens[idx])}


class BucketSampler:
    """dataset.py:35-119.  Yields lists of dataset indices."""

    def __init__(self, ilens, min_ilen, max_ilen, half_batch_ilen, batch_size, bucket_size, bucket_reverse, drop_last):
        self.ilens, self.min_ilen, self.max_ilen = ilens, min_ilen, max_ilen
        self.half_batch_ilen = half_batch_ilen if half_batch_ilen else ILEN_MAX
        self.batch_size, self.bucket_size = batch_size, bucket_size
        self.drop_last, self.bucket_reverse = drop_last, bucket_reverse
        lb = min(ILEN_MIN, bucket_size) if not min_ilen else min_ilen
        ub = max(ILEN_MAX, ilens.max()) if not max_ilen else max_ilen
        bins = np.arange(ub, lb, -bucket_size) if bucket_reverse else np.arange(lb, ub, bucket_size)
        bucket_idx = np.digitize(ilens, bins, right=True)
        self.half_batch_size_bucket_idx = np.digitize(self.half_batch_ilen, bins, right=True)
        self.buckets = []
        for bin_idx in range(1, len(bins) - 1):          # utterances outside (lb, ub) fall into no bucket
            bucket = np.where(bucket_idx == bin_idx)[0]
            if len(bucket) > 0:
                self.buckets.append((bin_idx, bucket))
        random.shuffle(self.buckets)

    def _get_batch_size(self, bin_idx):
        half = max(1, self.batch_size // 2)
        if self.bucket_reverse:
            return half if bin_idx < self.half_batch_size_bucket_idx else self.batch_size
        return half if bin_idx > self.half_batch_size_bucket_idx else self.batch_size

    def __iter__(self):
        for bin_idx, bucket in self.buckets:
            batch_size = self._get_batch_size(bin_idx)
            np.random.shuffle(bucket)
            batch = []
            for idx in bucket:
                batch.append(idx)
                if len(batch) == batch_size:
                    yield batch
                    batch = []
            if len(batch) > 0 and not self.drop_last:
                yield batch

    def __len__(self):
        n = 0
        for bin_idx, bucket in self.buckets:
            bs = self._get_batch_size(bin_idx)
            n += len(bucket) // bs if self.drop_last else (len(bucket) + bs - 1) // bs
        return n


class _PlainBatches:
    """The non-bucket path of get_loader: torch's Sequential / RandomSampler order in chunks of batch_size
    (what DataLoader(batch_size=, shuffle=) does), over `indices` (a random_split subset or everything)."""

    def __init__(self, indices, batch_size, shuffle, drop_last):
        self.indices, self.batch_size, self.shuffle, self.drop_last = indices, batch_size, shuffle, drop_last

    def __iter__(self):
        n = len(self.indices)
        if self.shuffle:
            order = list(torch.utils.data.RandomSampler(range(n)))      # consumes the torch RNG like the reference
        else:
            order = range(n)
        batch = []
        for i in order:
            batch.append(int(self.indices[i]))
            if len(batch) == self.batch_size:
                yield batch
                batch = []
        if batch and not self.drop_last:
            yield batch

    def __len__(self):
        n = len(self.indices)
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size


def assemble_batch(dset, idxs, pin=False):
    """collate_fn (dataset.py:21-33) without per-utterance tensors: stable descending sort by ilen, rows copied
    from the feature array into one zero-padded [B, Tmax, D] buffer.  Pinned buffers come from torch's caching host
    allocator: after the first epoch no cudaHostAlloc happens, and a block is not handed out again before the
    asynchronous host->device copies that read it have completed."""
    idxs = sorted(idxs, key=lambda i: int(dset.ilens[i]), reverse=True)       # list.sort is stable, like the reference
    il = np.asarray([dset.ilens[i] for i in idxs])
    B, Tmax, Dm = len(idxs), int(il.max()), int(dset.feat.shape[1])
    xs = torch.empty((B, Tmax, Dm), dtype=torch.float32, pin_memory=pin)
    xn = xs.numpy()
    for b, i in enumerate(idxs):
        n = int(il[b])
        xn[b, :n] = dset.feat[dset.iptr[i]:dset.iptr[i + 1], :]
        if n < Tmax:
            xn[b, n:] = 0.0
    ilens = torch.as_tensor(il)
    ys = [torch.as_tensor(dset.label[dset.optr[i]:dset.optr[i + 1]]) for i in idxs]
    olens = torch.as_tensor(np.asarray([dset.olens[i] for i in idxs]))
    return xs, ilens, ys, olens


class B200Loader:
    """Iterable of (xs_pad, ilens, ys, olens) like the reference's DataLoader; see the module docstring."""

    def __init__(self, dset, batches, prefetch=2, pin_memory=True, device=None):
        self.dset, self.batches, self.prefetch = dset, batches, max(0, int(prefetch))
        self.pin = bool(pin_memory) and torch.cuda.is_available()
        self.device = torch.device(device) if device is not None else None
        self.ilens = dset.ilens

    def __len__(self):
        return len(self.batches)

    def _load(self, idxs, stream):
        out = assemble_batch(self.dset, idxs, self.pin)
        if self.device is not None and self.device.type == "cuda":
            with torch.cuda.stream(stream):
                xd = out[0].to(self.device, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(stream)
            return (xd,) + out[1:], ev
        return out, None

    def __iter__(self):
        # torch's DataLoader iterator draws its base seed from the default torch generator when it is created
        # (before the sampler draws anything): keep the torch RNG stream aligned with the reference's
        torch.empty((), dtype=torch.int64).random_()
        return self._iterate()

    def _iterate(self):
        stream = torch.cuda.Stream(self.device) if (self.device is not None and self.device.type == "cuda") else None
        if self.prefetch == 0:
            for idxs in self.batches:
                out, ev = self._load(idxs, stream)
                if ev is not None:
                    cur = torch.cuda.current_stream(self.device)
                    cur.wait_event(ev)
                    out[0].record_stream(cur)
                yield out
            return
        # index lists are drawn HERE (consumer thread, reference order); the worker only moves bytes
        todo, done = queue.Queue(), queue.Queue()
        stop = object()

        def worker():
            while True:
                idxs = todo.get()
                if idxs is stop:
                    return
                try:
                    done.put(self._load(idxs, stream))
                except BaseException as e:               # surfaced to the consumer
                    done.put(e)
        th = threading.Thread(target=worker, daemon=True)
        th.start()
        it = iter(self.batches)
        inflight = 0
        try:
            for _ in range(self.prefetch):
                idxs = next(it, None)
                if idxs is None:
                    break
                todo.put(idxs)
                inflight += 1
            while inflight:
                res = done.get()
                inflight -= 1
                if isinstance(res, BaseException):
                    raise res
                idxs = next(it, None)
                if idxs is not None:
                    todo.put(idxs)
                    inflight += 1
                out, ev = res
                if ev is not None:
                    cur = torch.cuda.current_stream(self.device)
                    cur.wait_event(ev)
                    out[0].record_stream(cur)      # allocated on the loader's stream, consumed on the trainer's
                yield out
        finally:
            todo.put(stop)


def get_loader(data_dir, batch_size, is_memmap, is_bucket, num_workers=0, split_rate=1.0, split_seed=531,
               min_ilen=None, max_ilen=None, half_batch_ilen=None, bucket_reverse=False, shuffle=True, read_file=False,
               drop_last=False, pin_memory=True, device=None):
    """dataset.py:158-205.  `num_workers` is read as the prefetch depth (0 = synchronous, reference-exact RNG
    interleaving); `device` additionally stages the features on that CUDA device."""
    assert not read_file, "Load from Kaldi ark haven't been implemented yet"
    dset = CommonVoiceDataset(Path(data_dir), is_memmap)
    indices = np.arange(len(dset))
    if split_rate < 1.0:
        tot = len(dset)
        num_tr = int(tot * split_rate)
        sub, _ = torch.utils.data.random_split(range(tot), [num_tr, tot - num_tr],
                                               generator=torch.Generator().manual_seed(split_seed))
        indices = np.asarray(sub.indices)
    prefetch = 0 if not is_memmap else 2 * int(num_workers)
    if is_bucket and split_rate == 1.0:
        batches = BucketSampler(dset.ilens, min_ilen=min_ilen, max_ilen=max_ilen, half_batch_ilen=half_batch_ilen,
                                batch_size=batch_size, bucket_size=BUCKET_SIZE, bucket_reverse=bucket_reverse,
                                drop_last=drop_last)
    else:
        batches = _PlainBatches(indices, batch_size, shuffle, drop_last)
    return B200Loader(dset, batches, prefetch=prefetch, pin_memory=pin_memory, device=device)


class DataContainer:
    """dataset.py:207-277: one endless train iterator and one dev loader per accent."""

    def __init__(self, data_dirs, batch_size, dev_batch_size, is_memmap, is_bucket, num_workers=0, min_ilen=None,
                 max_ilen=None, half_batch_ilen=None, bucket_reverse=False, shuffle=True, read_file=False,
                 drop_last=False, pin_memory=True, device=None):
        self.data_dirs = [Path(d) for d in data_dirs]
        self.num_datasets = len(self.data_dirs)
        self.kw = dict(batch_size=batch_size, is_memmap=is_memmap, is_bucket=is_bucket, num_workers=num_workers,
                       min_ilen=min_ilen, max_ilen=max_ilen, half_batch_ilen=half_batch_ilen,
                       bucket_reverse=bucket_reverse, shuffle=shuffle, read_file=read_file, device=device)
        self.reload_cnt = 0
        self.loader_iters, self.dev_loaders = [], []
        for d in self.data_dirs:
            self.loader_iters.append(iter(get_loader(d.joinpath('train'), **self.kw)))
            self.dev_loaders.append(get_loader(d.joinpath('dev'), batch_size=dev_batch_size, is_memmap=is_memmap,
                                               is_bucket=False, num_workers=num_workers, shuffle=False, device=device))

    def get_item(self, accent_idx=None, num=1):
        ret_ls = []
        if accent_idx is None:                            # MultiASRInterface
            accent_ids = np.random.randint(self.num_datasets, size=num)
        else:
            accent_ids = np.repeat(accent_idx, num)
        for accent_id in accent_ids:
            try:
                ret = next(self.loader_iters[accent_id])
            except StopIteration:
                self.loader_iters[accent_id] = iter(get_loader(self.data_dirs[accent_id].joinpath('train'), **self.kw))
                self.reload_cnt += 1
                ret = next(self.loader_iters[accent_id])
            ret_ls.append((accent_id, ret))
        return ret_ls
