"""CPU tests: oracle/port.py (the checker that travels to the GPU box) against the golden
vectors produced by the live reference (oracle/make_golden.py).  This is what pins the
oracle; see SURVEY.md section 8(c)."""
import math

import numpy as np
import torch

from oracle import port
from tests.helpers import GOLD, check_adam_weights, check_summary, load_batch, load_weights, tiny_cfg


def test_masks_lengths_targets_bit_exact():
    z = np.load(GOLD / "run_batch_tiny.npz")
    cfg = tiny_cfg()
    x, ilens, ys, olens = load_batch(z, "in.")
    enc_lens = port.enc_lengths(ilens)
    assert np.array_equal(enc_lens.numpy(), z["enc_lens"])
    assert np.array_equal(port.make_bool_pad_mask(enc_lens).numpy(), z["enc_pad_mask"])
    L1 = z["gold"].shape[1]
    assert np.array_equal(port.generate_square_subsequent_mask(L1).numpy(), z["causal_mask"])
    ys_in, ys_out = port.prepare_targets(ys, cfg)
    assert np.array_equal(ys_out.numpy(), z["gold"])
    assert ys_in[:, 0].eq(0).all() and ys_in[1, 4:].eq(366).all()
    pe = port.positional_table(3000, cfg.d_model)
    assert np.array_equal(pe[:64, 0].numpy(), z["pe_head"])


def test_forward_loss_grads_match_reference():
    z = np.load(GOLD / "run_batch_tiny.npz")
    cfg = tiny_cfg()
    sd = load_weights(cfg)
    x, ilens, ys, olens = load_batch(z, "in.")
    info, grads, logit, gold = port.run_batch(sd, cfg, x, ilens, ys, olens, 0.2)
    assert np.array_equal(olens.numpy(), z["olens_after"])          # olens += 1 in place
    assert np.abs(logit.numpy() - z["logit"]).max() < 2e-5
    assert np.array_equal(logit.argmax(-1).numpy(), z["logit"].argmax(-1))
    assert abs(info["loss"] - float(z["loss"])) <= 1e-5 * abs(float(z["loss"]))
    assert info["acc"] == float(z["acc"])
    for n, g in grads.items():
        check_summary(z, "g.", n, g, rtol_l2=2e-3, atol_sample=2e-3)


def test_greedy_decode_ids_bit_exact():
    z = np.load(GOLD / "run_batch_tiny.npz")
    cfg = tiny_cfg()
    sd = load_weights(cfg)
    x, ilens, _, _ = load_batch(z, "in.")
    with torch.no_grad():
        ids = port.greedy_decode(sd, cfg, x, ilens)
    assert np.array_equal(ids.numpy(), z["greedy"])


def test_fomaml_meta_steps_match_reference():
    z = np.load(GOLD / "fomaml_tiny.npz")
    cfg = tiny_cfg()
    ml = port.MetaLearner(load_weights(cfg), cfg, algo="fomaml", k=float(z["k"]), warmup=int(z["warmup_steps"]))
    assert abs(ml.inner_lr - float(z["inner_lr"])) < 1e-15
    for step in range(int(z["n_meta_steps"])):
        for acc in range(int(z["n_accents"])):
            tr = [load_batch(z, f"s{step}.a{acc}.tr{j}.") for j in range(int(z["meta_k"]))]
            ml.run_task(tr)
            info = ml.inner_test_and_accumulate(load_batch(z, f"s{step}.a{acc}.te."))
            ref = float(z[f"s{step}.a{acc}.te_loss"])
            assert abs(info["loss"] - ref) <= 2e-4 * abs(ref), (step, acc, info["loss"], ref)
        lr = ml.final_meta_update()
        assert abs(lr - float(z[f"s{step}.lr"])) < 1e-12
        for n, g in ml.last_meta_grad.items():
            check_summary(z, f"s{step}.mg.", n, g, rtol_l2=5e-3, atol_sample=1e-2)
        for n in ml.meta_names:
            check_adam_weights(z, f"s{step}.w.", [f"s{i}.mg." for i in range(step + 1)], n,
                               ml.original[n], lr)
    # asr_model is left holding the last task's fast weights
    for n in ml.meta_names:
        check_adam_weights(z, "fast.", ["s0.mg."], n, ml.fast[n], lr)


def test_multi_steps_match_reference():
    z = np.load(GOLD / "multi_tiny.npz")
    cfg = tiny_cfg()
    ml = port.MetaLearner(load_weights(cfg), cfg, algo="multi", k=float(z["k"]), warmup=int(z["warmup_steps"]))
    for step in range(int(z["n_steps"])):
        info = ml.multi_step(load_batch(z, f"s{step}."))
        ref = float(z[f"s{step}.loss"])
        assert abs(info["loss"] - ref) <= 5e-4 * abs(ref), (step, info["loss"], ref)
        lr = port.noam_lr(step + 1, float(z["k"]), cfg.d_model, int(z["warmup_steps"]))
        for n in ml.names:
            check_adam_weights(z, f"s{step}.w.", [f"s{i}.g." for i in range(step + 1)], n, ml.fast[n], lr)


def test_reptile_definition():
    """Reptile has no reference (fo_meta_interface.py:197-198 raises): check the build-defined
    semantics: meta-gradient = mean over tasks of (theta - phi_task)."""
    z = np.load(GOLD / "fomaml_tiny.npz")
    cfg = tiny_cfg()
    ml = port.MetaLearner(load_weights(cfg), cfg, algo="reptile", k=1.0, warmup=4)
    theta = {n: t.clone() for n, t in ml.original.items()}
    phis = []
    for acc in range(2):
        ml.run_task([load_batch(z, f"s0.a{acc}.tr{j}.") for j in range(2)])
        phis.append({n: t.clone() for n, t in ml.fast.items()})
        ml.inner_test_and_accumulate(load_batch(z, f"s0.a{acc}.te."))
    ml.final_meta_update()
    for n in ml.meta_names:
        src = "char_trans.weight" if n == "pre_embed.weight" else n
        want = sum(theta[n] - p[src] for p in phis) / 2
        assert torch.allclose(ml.last_meta_grad[n], want, atol=1e-7)


def test_ctc_port_matches_reference_call_site():
    z = np.load(GOLD / "ctc.npz")
    for c in ("a", "b"):
        logits = torch.from_numpy(z[f"{c}.logits"].copy()).transpose(0, 1).contiguous()   # [T,B,C]
        nll, loss, grad = port.ctc_alpha_beta(
            logits, torch.from_numpy(z[f"{c}.targets"]), torch.from_numpy(z[f"{c}.in_lens"]),
            torch.from_numpy(z[f"{c}.tgt_lens"]))
        assert np.allclose(nll.numpy(), z[f"{c}.nll"], rtol=1e-5, atol=1e-5)
        assert abs(float(loss) - float(z[f"{c}.loss"])) <= 1e-5 * abs(float(z[f"{c}.loss"]))
        g_ref = z[f"{c}.grad_logits"]                                                  # [B,T,C]
        # the golden was computed in fp32 by ATen (nll ~ 200 -> ~3e-5 relative noise)
        assert np.abs(grad.transpose(0, 1).numpy() - g_ref).max() < 1e-4 * np.abs(g_ref).max()
        # zero pattern: frames beyond input length and infeasible utterances
        assert (grad.transpose(0, 1).numpy()[g_ref == 0] == 0).all()


def test_noam_lr():
    assert math.isclose(port.noam_lr(1, 1.0, 512, 25000), 1.118e-8, rel_tol=1e-3)
    assert math.isclose(port.inner_lr(1.0, 512, 25000), 2.795e-4, rel_tol=1e-3)


def test_port_ctc_blank_as_label_matches_aten():
    """ATen allows a target id equal to the blank index and compares only the extended labels for the skip
    transition (a blank-valued label after a different label may be skipped to): the oracle follows that."""
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(1)
    T, B, C = 12, 2, 8
    lg = torch.randn(T, B, C, generator=g, dtype=torch.float64)
    tg = torch.tensor([4, 0, 6, 0, 0, 4, 0, 3])
    il, tl = torch.tensor([T, T - 1]), torch.tensor([6, 2])
    l = lg.clone().requires_grad_(True)
    ref = F.ctc_loss(F.log_softmax(l, -1), tg, il, tl, blank=0, reduction='mean', zero_infinity=True)
    ref.backward()
    onll, oloss, ograd = port.ctc_alpha_beta(lg.float(), tg, il, tl)
    assert abs(float(oloss) - float(ref.detach())) <= 1e-6 * abs(float(ref.detach()))
    assert float((ograd - l.grad).abs().max()) <= 1e-6
