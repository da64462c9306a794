"""CPU tests of the HOST-SIDE orchestration (engine.py): kernel schedule, buffer reuse, arena
layout, accumulation flags.  The kernels are replaced by the torch test double of
tests/torch_backend.py; results are compared with the golden vectors of the live reference."""
import numpy as np
import torch

from metaasr_crossaccent_b200.engine import ArenaLayout, NetConfig, TransformerEngine, param_shapes
from oracle import port
from tests.helpers import GOLD, check_summary, load_batch, load_weights, tiny_cfg
from tests.torch_backend import TorchBackend

TINY = dict(idim=83, d_model=32, nheads=4, d_inner=64, enc_layers=2, dec_layers=2, odim=367, dropout=0.0, pos_dropout=0.0)


def make_engine():
    cfg = NetConfig(**TINY)
    eng = TransformerEngine(cfg, TorchBackend("cpu"), "cpu", label_smoothing=0.2)
    eng.load_state_dict(load_weights(tiny_cfg()))
    return eng


def test_param_shapes_match_oracle_and_reference_keys():
    cfg = NetConfig(**TINY)
    assert list(param_shapes(cfg).items()) == list(port.param_shapes(tiny_cfg()).items())
    big = NetConfig()
    shapes = param_shapes(big)
    assert len(shapes) == 114
    lay = ArenaLayout(big)
    assert len(lay.offsets) == 112 and lay.n_unique_elems == 24_881_455
    assert all(off % 64 == 0 for off in lay.offsets.values())


def test_engine_forward_backward_matches_reference():
    z = np.load(GOLD / "run_batch_tiny.npz")
    eng = make_engine()
    x, ilens, ys, olens = load_batch(z, "in.")
    hb = eng.prepare_batch(x, ilens, ys, olens)
    assert np.array_equal(olens.numpy(), z["olens_after"])
    assert np.array_equal(hb["ys_out"].numpy(), z["gold"])
    assert np.array_equal(hb["enc_lens"].numpy(), z["enc_lens"])
    db = eng.to_device(hb)
    ws = eng.forward_backward(db)
    logit = ws["logits"].view(3, -1, 367)
    assert np.abs(logit.numpy() - z["logit"]).max() < 2e-5
    info = eng.read_stats()
    assert abs(info["loss"] - float(z["loss"])) <= 1e-5 * abs(float(z["loss"]))
    assert info["acc"] == float(z["acc"])
    assert np.array_equal(ws["argmax"].view(3, -1).numpy(), z["logit"].argmax(-1))
    for n in eng.layout.offsets:
        check_summary(z, "g.", n, eng.G[n], rtol_l2=2e-3, atol_sample=2e-3)


def test_state_dict_layout_is_reference_compatible():
    eng = make_engine()
    sd = eng.state_dict()
    assert list(sd.keys()) == list(port.param_shapes(tiny_cfg()).keys())
    assert sd["char_trans.weight"].data_ptr() == sd["pre_embed.weight"].data_ptr()       # tied
    assert sd["pos_encoder.pe"].shape == (3000, 1, 32)
    assert all(t.dtype == torch.float32 for t in sd.values())


def test_joint_ctc_attention_objective_matches_torch_restatement():
    """ctc_weight extension (north_star kernel 3, parity unpinned: no reference): the engine's schedule with a CTC head
    on the encoder memory -- (1-w) LS-CE + w CTC, gradient scales folded into the producing kernels, the CTC gradient
    entering the encoder next to the decoder's -- against autograd on oracle/port.run_batch_joint; and w = 0 leaves the
    reference's 114-key state dict untouched."""
    w = 0.3
    cfg = NetConfig(**TINY, ctc_weight=w)
    assert list(param_shapes(cfg))[:-2] == list(param_shapes(NetConfig(**TINY))) and list(param_shapes(cfg))[-2:] == ["ctc_lo.weight", "ctc_lo.bias"]
    lay0, lay1 = ArenaLayout(NetConfig(**TINY)), ArenaLayout(cfg)
    assert all(lay1.offsets[n] == o for n, o in lay0.offsets.items())          # reference tensors keep their offsets
    eng = TransformerEngine(cfg, TorchBackend("cpu"), "cpu", label_smoothing=0.2)
    sd = load_weights(tiny_cfg())
    g = torch.Generator().manual_seed(3)
    sd["ctc_lo.weight"] = torch.randn(367, 32, generator=g) * 0.1
    sd["ctc_lo.bias"] = torch.randn(367, generator=g) * 0.1
    eng.load_state_dict(sd)
    z = np.load(GOLD / "run_batch_tiny.npz")
    x, ilens, ys, olens = load_batch(z, "in.")
    info_ref, grads, _, _ = port.run_batch_joint(sd, tiny_cfg(), x, ilens, ys, olens.clone(), 0.2, w, training=False)
    hb = eng.prepare_batch(x, ilens, ys, olens)
    ws = eng.forward(hb, want_grad=True)
    eng.backward(hb, ws)
    info = eng.read_stats()
    assert abs(info["loss"] - info_ref["loss"]) <= 1e-5 * abs(info_ref["loss"])
    assert abs(info["ctc_loss"] - info_ref["ctc_loss"]) <= 1e-5 * abs(info_ref["ctc_loss"]) and info["ctc_loss"] > 0
    assert abs(info["att_loss"] - info_ref["att_loss"]) <= 1e-5 * abs(info_ref["att_loss"])
    for n, gr in grads.items():
        mine = eng.G[n]
        rel = float((mine - gr).norm() / gr.norm().clamp_min(1e-12))
        assert rel < 2e-3, (n, rel)
