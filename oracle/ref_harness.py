"""TEST INFRASTRUCTURE (oracle side) -- never imported by the product package.

Runs the UNMODIFIED reference (sunprinceS/MetaASR-CrossAccent, mounted read-only
at /root/reference) on CPU so that its own classes produce golden vectors for the
hot path.  It only works in the build container: the GPU
box has no /root/reference, so everything produced here is committed as small
fixtures under tests/golden/ by oracle/make_golden.py.

What it does:
  * registers stub modules for the reference's missing third-party deps
    (tqdmlogger, torchexp.stat, torch_optimizer, comet_ml, editdistance, IPython);
  * neutralises nn.Module.cuda (the reference hard-codes .cuda(),
    src/transformer_torch_trainer.py:21);
  * builds a solver through the reference's own get_trainer(FOMetaASRInterface |
    MultiASRInterface, ...) from a scratch cwd that holds data/ symlinks.
"""
from __future__ import annotations

import argparse
import os
import sys
import tempfile
import types
from pathlib import Path

import torch

REFERENCE_ROOT = Path(os.environ.get("METAASR_REFERENCE", "/root/reference"))


def reference_available() -> bool:
    return (REFERENCE_ROOT / "src" / "fo_meta_interface.py").exists()


# --------------------------------------------------------------------------- stubs
class RunningAvgDict(dict):
    """Stand-in for torchexp.stat.RunningAvgDict (used at pretrain_interface.py:12).

    decay_rate == 1.0 -> count-weighted mean; otherwise exponential moving average.
    """

    def __init__(self, decay_rate=0.99):
        super().__init__()
        self.decay_rate = decay_rate
        self._n = {}

    def add(self, info, n=1):
        for k, v in info.items():
            v = float(v)
            if k not in self:
                self[k] = v
                self._n[k] = n
            elif self.decay_rate >= 1.0:
                tot = self._n[k] + n
                self[k] = (self[k] * self._n[k] + v * n) / tot
                self._n[k] = tot
            else:
                self[k] = self.decay_rate * self[k] + (1 - self.decay_rate) * v


def _levenshtein(a, b):
    a, b = list(a), list(b)
    prev = list(range(len(b) + 1))
    for i, ca in enumerate(a, 1):
        cur = [i]
        for j, cb in enumerate(b, 1):
            cur.append(min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (ca != cb)))
        prev = cur
    return prev[-1]


def install_stubs():
    if "tqdmlogger" in sys.modules:
        return

    tl = types.ModuleType("tqdmlogger")
    tl.log = lambda *a, **k: None
    tl.seclog = lambda *a, **k: None
    tl.flush = lambda *a, **k: None
    tl.logger = types.SimpleNamespace(info=lambda *a, **k: None)
    ansi = types.ModuleType("tqdmlogger.ansistyle")
    ansi.stylize = lambda s, *a: s
    ansi.fg = ansi.bg = ansi.attr = lambda c: ""
    ansi.RESET = ""
    tl.ansistyle = ansi
    sys.modules["tqdmlogger"] = tl
    sys.modules["tqdmlogger.ansistyle"] = ansi

    te = types.ModuleType("torchexp")
    tes = types.ModuleType("torchexp.stat")
    tes.RunningAvgDict = RunningAvgDict
    te.stat = tes
    sys.modules["torchexp"] = te
    sys.modules["torchexp.stat"] = tes

    sys.modules["torch_optimizer"] = types.ModuleType("torch_optimizer")

    class _Exp:
        alive = True

        def __init__(self, *a, **k):
            pass

        def get_key(self):
            return "oracle"

        def __getattr__(self, name):
            return lambda *a, **k: None

    cm = types.ModuleType("comet_ml")
    cm.Experiment = _Exp
    cm.ExistingExperiment = _Exp
    sys.modules["comet_ml"] = cm

    ed = types.ModuleType("editdistance")
    ed.eval = _levenshtein
    sys.modules["editdistance"] = ed

    ip = types.ModuleType("IPython")
    ip.embed = lambda *a, **k: None
    sys.modules["IPython"] = ip

    torch.nn.Module.cuda = lambda self, *a, **k: self
    if str(REFERENCE_ROOT) not in sys.path:
        sys.path.insert(0, str(REFERENCE_ROOT))


# --------------------------------------------------------------------------- solver
def make_workdir() -> Path:
    """Scratch cwd with data/ symlinks (YAML paths are cwd-relative)."""
    wd = Path(tempfile.mkdtemp(prefix="metaasr_oracle_"))
    (wd / "data").mkdir()
    for f in ("accent-code.json", "valid_train_en_unigram150.model",
              "valid_train_en_unigram150_units.txt"):
        os.symlink(REFERENCE_ROOT / "data" / f, wd / "data" / f)
    return wd


def base_config(d_model=512, nheads=8, d_inner=2048, enc_layers=2, dec_layers=4,
                dropout=0.0, warmup_steps=25000, k=1.0, label_smoothing=0.2,
                meta=True):
    """In-memory equivalent of config/transformer/pretrain/{fometa,multi}-hkust.yaml."""
    am = {
        "idim": 83, "nheads": nheads, "d_model": d_model, "d_inner": d_inner,
        "dropout": dropout, "tgt_share_weight": 1,
        "encoder": {"nlayers": enc_layers}, "decoder": {"nlayers": dec_layers},
        "pos_dropout": dropout,
    }
    if meta:
        am.update({
            "inner_optimizer_cls": "SGD",
            "inner_optimizer_opt": {"momentum": 0.9, "nesterov": True},
            "meta_opt_cls": "noam",
            "meta": {"optimizer_opt": {"k": k, "warmup_steps": warmup_steps}},
        })
    else:
        am.update({"optimizer_cls": "noam",
                   "optimizer_opt": {"k": k, "warmup_steps": warmup_steps}})
    solver = {
        "setting": "oracle", "data_root": "data", "total_steps": 1000000,
        "spm_mapping": "data/valid_train_en_unigram150_units.txt",
        "spm_model": "data/valid_train_en_unigram150.model",
        "label_smoothing": label_smoothing,
        "eval_ival": 5000, "log_ival": 1000000, "save_ival": 1000000,
        "batch_size": 32, "dev_batch_size": 32, "min_ilen": 10, "max_ilen": 1500,
        "dev_max_ilen": 3000, "half_batch_ilen": 512,
    }
    return {"asr_model": am, "solver": solver}


def make_paras(algo="fomaml", accents=("ca", "en"), meta_k=1, seed=531):
    return argparse.Namespace(
        config="<memory>", pretrain_suffix="oracle", pretrain_accents=list(accents),
        num_pretrain=len(accents), tgt_accent="hk", runs=0, overwrite=True, seed=seed,
        no_cuda=True, no_memmap=False, no_bucket=False, meta_k=meta_k,
        meta_batch_size=len(accents), sample_strategy="normal", max_step=0,
        resume=False, resume_step=-1, use_tensorboard=False, model_name="transformer",
        algo=algo, njobs=0, cuda=False, is_bucket=True, is_memmap=True)


def build_reference_solver(config, paras, seed=531):
    """get_trainer(InterfaceCls, config, paras, id2accent) exactly as pretrain.py:70-86,
    minus load_data() (batches are injected by the caller)."""
    import json
    import random
    import numpy as np

    install_stubs()
    wd = make_workdir()
    os.chdir(wd)
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    with open(Path("data", "accent-code.json")) as fin:
        id2accent = json.load(fin)
    if paras.algo == "multi":
        from src.multi_interface import MultiASRInterface as Iface
    else:
        from src.fo_meta_interface import FOMetaASRInterface as Iface
    from src.transformer_torch_trainer import get_trainer
    solver = get_trainer(Iface, config, paras, id2accent)
    solver.id2ch = solver.id2units       # what load_data() would do (pretrain_interface.py:111)
    solver.set_model()
    return solver
