"""Boundary proof: `get_trainer` of this repo takes the REFERENCE's own interface classes
(`src.fo_meta_interface.FOMetaASRInterface`, `src.multi_interface.MultiASRInterface`) and the reference's own YAML file
unchanged -- only `run_batch` / the model are replaced; `run_task`, `clip_grad_norm_`, `torch.optim.SGD`,
`_partial_meta_update`, `_final_meta_update` and the noam optimizer are the reference's code operating on the `.grad`s the
engine leaves on `asr_model.parameters()`.  CPU host logic on the torch test double; needs /root/reference (build
container only: the GPU box has no reference checkout, the test skips there)."""
import os
from collections import OrderedDict

import numpy as np
import pytest
import torch
import yaml

from oracle import port, ref_harness
from tests.helpers import clone_batch, hkust_profile_batch
from tests.torch_backend import TorchBackend

pytestmark = pytest.mark.skipif(not ref_harness.reference_available(), reason="reference checkout not present")


def _solver(algo, yaml_name):
    import json
    import random
    ref_harness.install_stubs()
    cwd = os.getcwd()
    os.chdir(ref_harness.make_workdir())
    try:
        with open(ref_harness.REFERENCE_ROOT / "config" / "transformer" / "pretrain" / yaml_name) as f:
            config = yaml.safe_load(f)                       # the reference's file, as pretrain.py:52 reads it
        config["asr_model"]["dropout"] = config["asr_model"]["pos_dropout"] = 0.0     # deterministic comparison
        paras = ref_harness.make_paras(algo, accents=("ca", "en"), meta_k=1)
        paras.backend_factory = lambda dtype: TorchBackend("cpu", dtype)
        random.seed(531); np.random.seed(531); torch.manual_seed(531)
        with open(os.path.join("data", "accent-code.json")) as fin:
            id2accent = json.load(fin)
        if algo == "multi":
            from src.multi_interface import MultiASRInterface as Iface
        else:
            from src.fo_meta_interface import FOMetaASRInterface as Iface
        from metaasr_crossaccent_b200.trainer import get_trainer          # <- the one import INTEGRATION.md changes
        s = get_trainer(Iface, config, paras, id2accent)
        s.id2ch = s.id2units                                             # what load_data() does (pretrain_interface.py:111)
        s.set_model()
    finally:
        os.chdir(cwd)
    return s, config


def _load_port_weights(s):
    cfg = port.NetCfg()
    sd = port.init_state_dict(cfg, seed=7)
    s.asr_model.load_state_dict(sd)
    return cfg, sd


@pytest.mark.timeout(600)
def test_reference_fomaml_interface_class_and_yaml_drive_the_engine():
    from torch import nn
    s, config = _solver("fomaml", "fometa-hkust.yaml")
    assert type(s).__mro__[1].__module__ == "src.fo_meta_interface"          # the reference's class, not ours
    assert len(s.asr_model.state_dict()) == 114
    cfg, sd = _load_port_weights(s)
    from src.nets_utils import clone_state_dict
    s._original = clone_state_dict(s.asr_model.state_dict(keep_vars=True))   # fo_meta_interface.py:100
    opt = config["asr_model"]["meta"]["optimizer_opt"]
    from src.model.transformer_pytorch.optimizer import TransformerOptimizer
    s.meta_opt = TransformerOptimizer(torch.optim.Adam(s._original.values(), betas=(0.9, 0.98), eps=1e-09), opt["k"],
                                      config["asr_model"]["d_model"], opt["warmup_steps"])   # :103-111
    tr = hkust_profile_batch(3, "eq", B=2, T=64, L=4)
    te = hkust_profile_batch(4, "eq", B=2, T=64, L=4)
    # the reference's own loop body (fo_meta_interface.py:139-156) on the reference's own methods
    s.run_task([(0, clone_batch(tr))])
    info = s._train(0, *clone_batch(te), accent_idx=0)
    gn = nn.utils.clip_grad_norm_(s.asr_model.parameters(), 5)
    assert np.isfinite(float(gn)) and set(info) == {"loss", "acc"}
    s._partial_meta_update()
    s._final_meta_update()
    # the same meta-step on the oracle
    ml = port.MetaLearner(sd, cfg, algo="fomaml", k=opt["k"], warmup=opt["warmup_steps"], eps_ls=0.2, training=False)
    infos, lr = ml.meta_step([([clone_batch(tr)], clone_batch(te))])
    assert abs(info["loss"] - infos[0]["loss"]) <= 1e-5 * abs(infos[0]["loss"])
    assert abs(s.meta_opt.lr - lr) < 1e-15 if hasattr(s.meta_opt, "lr") else True
    worst = 0.0
    for n in ml.meta_names:
        g = ml.last_meta_grad[n]
        mask = g.abs() > max(1e-7, 0.05 * float(g.abs().max()))
        err = (s._original[n].detach() - ml.original[n]).abs()[mask]
        if err.numel():
            worst = max(worst, float(err.max()) / lr)
    assert worst <= 3e-2, worst                        # post-Adam meta weights, in units of lr (see helpers.check_adam_weights)


@pytest.mark.timeout(600)
def test_reference_multi_interface_class_and_yaml_drive_the_engine():
    from torch import nn
    s, config = _solver("multi", "multi-hkust.yaml")
    assert type(s).__mro__[1].__module__ == "src.multi_interface"
    cfg, sd = _load_port_weights(s)
    b = hkust_profile_batch(5, "eq", B=2, T=64, L=4)
    # multi_interface.py:100-114 with the reference's objects: run_batch (ours) -> clip_grad_norm_ -> asr_opt.step()
    info = s._train(0, *clone_batch(b), accent_idx=0)
    gn = nn.utils.clip_grad_norm_(s.asr_model.parameters(), 5)
    assert np.isfinite(float(gn))
    before = s.asr_model.engine.params.clone()
    s.asr_opt.step()
    assert float((s.asr_model.engine.params - before).abs().max()) > 0
    oo = config["asr_model"]["optimizer_opt"]
    ml = port.MetaLearner(sd, cfg, algo="fomaml", k=oo["k"], warmup=oo["warmup_steps"], eps_ls=0.2, training=False)
    oinfo = ml.multi_step(clone_batch(b))
    assert abs(info["loss"] - oinfo["loss"]) <= 1e-5 * abs(oinfo["loss"])
    lr = port.noam_lr(1, oo["k"], cfg.d_model, oo["warmup_steps"])
    d = (s.asr_model.engine.P["encoder.layers.0.linear1.weight"] - ml.fast["encoder.layers.0.linear1.weight"]).abs()
    assert float(d.max()) <= 2.0 * lr + 1e-12
