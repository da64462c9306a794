"""GPU: CUDA-graph replay of run_batch (engine.use_graphs) is equivalent to launching every kernel from
the host -- same loss, same gradients, inputs re-loaded into the static buffers, fresh dropout masks."""
import numpy as np
import pytest
import torch

from tests.helpers import GOLD, load_batch
from tests.test_e2e_gpu import load_tiny, make_solver

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    return torch.device("cuda", 0)


@pytest.mark.parametrize("dtype,gemm", [("fp32", "simt"), ("bf16", "umma")])
def test_graph_replay_matches_eager(dev, dtype, gemm):
    z = np.load(GOLD / "run_batch_tiny.npz")
    s = make_solver("fomaml", dtype=dtype, gemm=gemm)
    load_tiny(s)
    eng = s.asr_model.engine
    b1 = lambda: load_batch(z, "in.")
    info_e = s.run_batch(0, *b1(), train=True)
    g_e = eng.grads.clone()
    eng.use_graphs = True
    info_g1 = s.run_batch(0, *b1(), train=True)          # capture + first replay
    info_g2 = s.run_batch(0, *b1(), train=True)          # pure replay
    assert abs(info_g1["loss"] - info_e["loss"]) <= 1e-6 * abs(info_e["loss"])
    assert info_g2["loss"] == info_g1["loss"]
    assert float((eng.grads - g_e).abs().max()) <= 1e-5 * float(g_e.abs().max())
    # different inputs through the same graph
    x, ilens, ys, olens = b1()
    x2 = x * 0.5
    info_g3 = s.run_batch(0, x2, ilens, ys, olens, train=True)
    eng.use_graphs = False
    x, ilens, ys, olens = b1()
    info_e3 = s.run_batch(0, x * 0.5, ilens, ys, olens, train=True)
    assert abs(info_g3["loss"] - info_e3["loss"]) <= 1e-6 * abs(info_e3["loss"])
    assert info_g3["loss"] != info_g1["loss"]


def test_graph_replay_draws_fresh_dropout_masks(dev):
    z = np.load(GOLD / "run_batch_tiny.npz")
    from tests.test_e2e_gpu import make_config
    import argparse
    from metaasr_crossaccent_b200 import interfaces as I
    from metaasr_crossaccent_b200.trainer import get_trainer
    cfg = make_config(True, dtype="fp32")
    cfg["asr_model"]["dropout"] = 0.3
    cfg["asr_model"]["pos_dropout"] = 0.3
    cfg["asr_model"]["cuda_graphs"] = True
    paras = argparse.Namespace(pretrain_accents=["ca", "en"], num_pretrain=2, tgt_accent="hk", runs=0, seed=531,
                               meta_k=2, meta_batch_size=2, sample_strategy="normal", max_step=0, resume=False,
                               algo="fomaml", pretrain_suffix="t", log_root=None)
    s = get_trainer(I.FOMetaASRInterface, cfg, paras, {"ca": "canada", "en": "england", "hk": "hongkong"})
    s.set_model()
    load_tiny(s)
    losses = [s.run_batch(0, *load_batch(z, "in."), train=True)["loss"] for _ in range(4)]
    assert len(set(losses)) == 4, losses                  # every replay sees a different mask
    assert max(losses) - min(losses) < 1.0
