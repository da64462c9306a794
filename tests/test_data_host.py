"""Input pipeline (metaasr_crossaccent_b200/data.py) against the batch sequences of the reference's own
src/io/dataset.py (tests/golden/loader.npz, made by oracle/make_golden.py golden_loader with num_workers=0):
same dataset indices in the same order in every batch (bucket sampler incl. half batches and the max_ilen cut,
sequential dev loader, torch-RandomSampler loader, DataContainer.get_item with epoch wrap-around and the
multi-task accent draw), same padded features / lengths / labels; the prefetching variant yields the same batches."""
import random

import numpy as np
import pytest
import torch

from metaasr_crossaccent_b200.data import DataContainer, get_loader
from tests.helpers import GOLD, make_synth_accent_dir


@pytest.fixture(scope="module")
def dirs(tmp_path_factory):
    root = tmp_path_factory.mktemp("accents")
    return [make_synth_accent_dir(root / f"acc{a}", seed=100 + a) for a in range(2)]


def check(z, prefix, batches):
    assert len(batches) == int(z[prefix + "n"]), prefix
    for i, (x, ilens, ys, olens) in enumerate(batches):
        assert np.array_equal(x[:, 0, 0].numpy().astype(np.int64), z[f"{prefix}{i}.idx"]), (prefix, i)
        assert np.array_equal(ilens.numpy(), z[f"{prefix}{i}.ilens"]) and np.array_equal(olens.numpy(), z[f"{prefix}{i}.olens"])
        assert tuple(x.shape) == tuple(z[f"{prefix}{i}.shape"])
        assert float(x.double().sum()) == float(z[f"{prefix}{i}.xsum"])          # same values, same zero padding
        assert sum(int(y.sum()) for y in ys) == int(z[f"{prefix}{i}.ysum"])
        assert [int(y.numel()) for y in ys] == olens.tolist() and x.dtype == torch.float32
        for b in range(x.shape[0]):
            assert float(x[b, int(ilens[b]):].abs().sum()) == 0.0


def seed(n):
    random.seed(n); np.random.seed(n); torch.manual_seed(n)


def test_bucket_dev_and_shuffled_loaders_match_reference(dirs):
    z = np.load(GOLD / "loader.npz")
    seed(7)
    ld = get_loader(dirs[0] / "train", batch_size=8, is_memmap=True, is_bucket=True, num_workers=0, min_ilen=None,
                    max_ilen=70, half_batch_ilen=50)
    assert len(ld) == int(z["bucket.e0.n"])
    check(z, "bucket.e0.", [tuple(t if not torch.is_tensor(t) else t.clone() for t in b) for b in ld])
    check(z, "bucket.e1.", [tuple(t if not torch.is_tensor(t) else t.clone() for t in b) for b in ld])
    check(z, "dev.", [tuple(t if not torch.is_tensor(t) else t.clone() for t in b)
                      for b in get_loader(dirs[0] / "dev", batch_size=5, is_memmap=True, is_bucket=False, shuffle=False)])
    seed(9)
    check(z, "shuf.", [tuple(t if not torch.is_tensor(t) else t.clone() for t in b)
                       for b in get_loader(dirs[1] / "train", batch_size=16, is_memmap=True, is_bucket=False, shuffle=True)])


def test_data_container_get_item_matches_reference(dirs):
    z = np.load(GOLD / "loader.npz")
    seed(11)
    dc = DataContainer(dirs, batch_size=8, dev_batch_size=5, is_memmap=True, is_bucket=True, num_workers=0,
                       max_ilen=70, half_batch_ilen=50)
    seq = [dc.get_item(accent_idx=a % 2, num=1)[0] for a in range(60)]
    seq += [it for _ in range(10) for it in dc.get_item(num=2)]
    assert np.array_equal(np.array([a for a, _ in seq]), z["dc.accents"]) and dc.reload_cnt == int(z["dc.reload_cnt"])
    check(z, "dc.", [tuple(t if not torch.is_tensor(t) else t.clone() for t in b) for _, b in seq])


def test_prefetching_loader_yields_the_same_batches(dirs):
    """num_workers > 0 = background assembly `2 * num_workers` batches ahead: same batches as the synchronous loader
    when nothing else draws from the numpy RNG in between (ring slots are not reused within the window)."""
    seed(21)
    a = [tuple(t if not torch.is_tensor(t) else t.clone() for t in b)
         for b in get_loader(dirs[0] / "train", batch_size=8, is_memmap=True, is_bucket=True, num_workers=0, half_batch_ilen=50)]
    seed(21)
    b = [tuple(t if not torch.is_tensor(t) else t.clone() for t in bb)
         for bb in get_loader(dirs[0] / "train", batch_size=8, is_memmap=True, is_bucket=True, num_workers=2, half_batch_ilen=50)]
    assert len(a) == len(b) > 10
    for (x1, i1, y1, o1), (x2, i2, y2, o2) in zip(a, b):
        assert torch.equal(x1, x2) and torch.equal(i1, i2) and torch.equal(o1, o2)
        assert all(torch.equal(u, v) for u, v in zip(y1, y2))


def test_pretrain_then_finetune_from_files(dirs, tmp_path):
    """The whole user-facing flow of pretrain.py / train.py on files in the reference's format, with the CPU test
    double as kernel backend: load_data() -> set_model() -> exec() for FOMAML (3 meta-steps, snapshot written), then
    the fine-tune loop that loads encoder modules from that snapshot and runs one epoch."""
    from tests.torch_backend import TorchBackend
    run_files_flow(dirs, tmp_path, lambda dt: TorchBackend("cpu", dt))


@pytest.mark.gpu
def test_pretrain_then_finetune_from_files_cuda(dirs, tmp_path):
    """Same flow on the CUDA path (C ABI kernels, pinned loader buffers, scorer fed by the CE kernel's argmax)."""
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    run_files_flow(dirs, tmp_path, None)


def run_files_flow(dirs, tmp_path, backend_factory):
    import argparse
    from metaasr_crossaccent_b200 import interfaces as I
    from metaasr_crossaccent_b200.trainer import get_trainer
    root = dirs[0].parent
    id2accent = {"a0": "acc0", "a1": "acc1"}
    mapping = tmp_path / "units.txt"           # a real unit inventory switches the CER / WER scorer on (metric.Metric)
    mapping.write_text("".join(f"{'▁' if i % 3 == 0 else ''}{chr(0x61 + i % 26)}{i % 7} {i}\n" for i in range(1, 366)))
    am = {"idim": 83, "nheads": 4, "d_model": 32, "d_inner": 64, "dropout": 0.0, "tgt_share_weight": 1,
          "encoder": {"nlayers": 1}, "decoder": {"nlayers": 1}, "pos_dropout": 0.0,
          "inner_optimizer_cls": "SGD", "inner_optimizer_opt": {"momentum": 0.9, "nesterov": True},
          "meta_opt_cls": "noam", "meta": {"optimizer_opt": {"k": 0.02, "warmup_steps": 4}}}
    solver = {"setting": "t", "total_steps": 4, "label_smoothing": 0.2, "eval_ival": 3, "log_ival": 1, "save_ival": 3,
              "data_root": str(root), "batch_size": 4, "dev_batch_size": 4, "min_ilen": None, "max_ilen": 70,
              "half_batch_ilen": 50, "spm_mapping": str(mapping), "spm_model": "/nonexistent.model",
              "dev_max_ilen": 1000}
    paras = argparse.Namespace(pretrain_accents=["a0", "a1"], num_pretrain=2, tgt_accent="a1", runs=0, seed=531, meta_k=2,
                               meta_batch_size=2, sample_strategy="normal", max_step=0, resume=False, algo="fomaml",
                               pretrain_suffix="t", log_root=str(tmp_path), is_memmap=True, is_bucket=True, njobs=1,
                               backend_factory=backend_factory)
    seed(5)
    s = get_trainer(I.FOMetaASRInterface, {"asr_model": am, "solver": solver}, paras, id2accent)
    s.load_data(); s.set_model()
    w0 = s._original_flat.clone()
    s.exec()
    assert s.global_step == 4 and not torch.equal(w0, s._original_flat)
    snap = s.log_dir / "snapshot.step.3"
    assert snap.exists() and s.data_container.num_datasets == 2
    # evaluate() ran at step 3: per-accent and averaged dev logs, scored hypotheses, best-model bookkeeping
    for f in ("dev_acc0_loss", "dev_acc1_cer", "dev_avg_wer", "best_wer", "best_cer", "model.wer.best", "train_loss"):
        assert (s.log_dir / f).exists(), f
    assert 0.0 < s.best_cer < 200.0 and 0.0 < s.best_wer < 200.0
    # fine-tune on accent a1 from that snapshot
    am2 = {k: v for k, v in am.items() if not k.startswith(("inner_", "meta"))}
    am2.update({"optimizer_cls": "noam", "optimizer_opt": {"k": 0.02, "warmup_steps": 4}})
    solver2 = dict(solver, total_epochs=1, eval_ival=1000, pretrain_module=["feat_extractor", "vgg2enc", "encoder"],
                   dev_max_ilen=1000)
    p2 = argparse.Namespace(accent="a1", runs=0, seed=531, algo="fomaml", pretrain=True, pretrain_model_path=str(snap),
                            pretrain_suffix="t", eval_suffix="ft", resume=False, save_verbose=False, eval_every_epoch=False,
                            log_root=str(tmp_path), is_memmap=True, is_bucket=True, njobs=0,
                            backend_factory=backend_factory)
    seed(6)
    f = get_trainer(I.MonoASRInterface, {"asr_model": am2, "solver": solver2}, p2, id2accent)
    f.load_data(); f.set_model()
    pre = torch.load(snap)
    assert torch.equal(f.asr_model.state_dict()["encoder.layers.0.linear1.weight"].cpu(), pre["encoder.layers.0.linear1.weight"].cpu())
    n_batches = len(f.train_set)
    f.exec()
    assert f.ep == 1 and f.global_step == 1 + n_batches and (f.log_dir / "snapshot.latest").exists()
    assert (f.log_dir / "dev_loss").exists() and (f.log_dir / "dev_cer").exists() and (f.log_dir / "model.wer.best").exists()


@pytest.mark.gpu
def test_device_staging_loader_matches_host_loader_and_feeds_run_batch(dirs):
    """get_loader(device=...): the features are staged on the GPU by the loader's own stream (with and without the
    background thread); same batches as the host loader, and run_batch takes the resident features as they are."""
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    import argparse
    from metaasr_crossaccent_b200 import interfaces as I
    from metaasr_crossaccent_b200.trainer import get_trainer
    kw = dict(batch_size=8, is_memmap=True, is_bucket=True, half_batch_ilen=50)
    seed(31)
    host = [tuple(t if not torch.is_tensor(t) else t.clone() for t in b) for b in get_loader(dirs[0] / "train", num_workers=0, **kw)]
    for nw in (0, 2):
        seed(31)
        devb = list(get_loader(dirs[0] / "train", num_workers=nw, device="cuda:0", **kw))
        assert len(devb) == len(host)
        for (x1, i1, y1, o1), (x2, i2, y2, o2) in zip(host, devb):
            assert x2.is_cuda and torch.equal(x1, x2.cpu()) and torch.equal(i1, i2) and torch.equal(o1, o2)
            assert all(torch.equal(u, v) for u, v in zip(y1, y2))
    am = {"idim": 83, "nheads": 4, "d_model": 32, "d_inner": 64, "dropout": 0.0, "tgt_share_weight": 1,
          "encoder": {"nlayers": 1}, "decoder": {"nlayers": 1}, "pos_dropout": 0.0,
          "optimizer_cls": "noam", "optimizer_opt": {"k": 0.02, "warmup_steps": 4}}
    solver = {"setting": "t", "total_steps": 4, "label_smoothing": 0.2, "eval_ival": 100, "log_ival": 100, "save_ival": 100}
    paras = argparse.Namespace(pretrain_accents=["a0"], num_pretrain=1, tgt_accent="a0", runs=0, seed=531, max_step=0,
                               resume=False, algo="multi", pretrain_suffix="t", log_root=None)
    s = get_trainer(I.MultiASRInterface, {"asr_model": am, "solver": solver}, paras, {"a0": "acc0"})
    s.set_model()
    sd = {k: v.clone() for k, v in s.asr_model.state_dict().items()}
    xh, ih, yh, oh = host[0]
    a = s.run_batch(0, xh, ih, [y.clone() for y in yh], oh.clone(), train=True)
    s.asr_model.load_state_dict(sd)
    xd, idv, yd, od = devb[0]
    b = s.run_batch(0, xd, idv, [y.clone() for y in yd], od.clone(), train=True)
    assert a["loss"] == b["loss"] and a["acc"] == b["acc"]
