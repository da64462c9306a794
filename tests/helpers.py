"""Shared helpers for the parity tests (oracle side; test infrastructure)."""
from __future__ import annotations

from collections import OrderedDict
from pathlib import Path

import numpy as np
import torch

from oracle import port

ROOT = Path(__file__).resolve().parents[1]
GOLD = ROOT / "tests" / "golden"
TINY = dict(d_model=32, nheads=4, d_inner=64, enc_layers=2, dec_layers=2)


def tiny_cfg(**kw):
    return port.NetCfg(idim=83, odim=367, **{**TINY, **kw})


def load_weights(cfg, dtype=torch.float32):
    w = np.load(GOLD / "weights_tiny.npz")
    sd = OrderedDict()
    for n in port.param_shapes(cfg):
        if n == "pos_encoder.pe":
            sd[n] = port.positional_table(3000, cfg.d_model, dtype)
        else:
            sd[n] = torch.from_numpy(w[n].copy()).to(dtype)
    if cfg.tie:
        sd["char_trans.weight"] = sd["pre_embed.weight"]
    return sd


def load_batch(z, prefix):
    x = torch.from_numpy(z[prefix + "x"].copy())
    ilens = torch.from_numpy(z[prefix + "ilens"].copy())
    olens = torch.from_numpy(z[prefix + "olens"].copy())
    cat = torch.from_numpy(z[prefix + "ys_cat"].copy())
    ys, o = [], 0
    for l in olens.tolist():
        ys.append(cat[o:o + l].clone())
        o += l
    return x, ilens, ys, olens


def summary(t, nsample=384):
    a = t.detach().cpu().double().numpy().reshape(-1)
    stride = max(1, a.size // nsample)
    return a[::stride][:nsample], a.sum(), np.sqrt((a ** 2).sum())


def check_summary(z, prefix, name, t, rtol_l2, atol_sample):
    """Compare a tensor with its golden (strided sample, sum, L2)."""
    s, _, l2 = summary(t)
    gs, gl2 = z[prefix + name + "#sample"].astype(np.float64), float(z[prefix + name + "#l2"])
    assert s.shape == gs.shape, (name, s.shape, gs.shape)
    scale = max(np.abs(gs).max(), 1e-30)
    err = np.abs(s - gs).max()
    assert err <= atol_sample * scale + 1e-30, f"{prefix}{name}: sample err {err:.3e} (scale {scale:.3e})"
    assert abs(l2 - gl2) <= rtol_l2 * max(gl2, 1e-30) + 1e-30, f"{prefix}{name}: l2 {l2} vs {gl2}"


def check_adam_weights(z, wprefix, gprefixes, name, t, lr, tol_lr=3e-2, gmin=1e-7):
    """Post-Adam parameters: Adam(eps=1e-9) turns a gradient of magnitude <~1e-8 (e.g. the key
    bias of every attention block, whose true gradient is 0) into a +-lr step of arbitrary sign,
    which not even the reference reproduces under fp32 re-ordering.  Compare the
    sampled elements whose golden gradient exceeded gmin in every step so far, to tol_lr * lr."""
    s, _, _ = summary(t)
    gs = z[wprefix + name + "#sample"].astype(np.float64)
    mask = np.ones_like(gs, dtype=bool)
    for gp in gprefixes:
        ga = np.abs(z[gp + name + "#sample"])
        # a gradient below the fp32 re-ordering noise of its tensor (~1 % of the largest entry after
        # two inner SGD steps through ReLU / max-pool switches) can flip sign -> +-2 lr under Adam
        mask &= ga > max(gmin, 0.05 * float(ga.max()))
    err = np.abs(s - gs)[mask]
    assert mask.sum() == 0 or err.max() <= tol_lr * lr, \
        f"{wprefix}{name}: max err {err.max():.3e} = {err.max() / lr:.3f} lr over {mask.sum()} elems"


def mono_paras(tmp_path, pre_path, **kw):
    """argparse.Namespace of train.py for the fine-tune loop (pretrained snapshot given by path)."""
    import argparse
    d = dict(accent="hk", runs=0, seed=531, algo="fomaml", pretrain=True, pretrain_model_path=str(pre_path),
             pretrain_suffix="t", eval_suffix="ft", resume=False, overwrite=True, save_verbose=False,
             eval_every_epoch=False, log_root=str(tmp_path), model_name="transformer")
    d.update(kw)
    return argparse.Namespace(**d)


def run_mono_freeze_check(make_solver_fn, z, tmp_path, loss_rtol, tol_lr=3e-2):
    """Shared by the host-logic (CPU double) and CUDA tests: the fine-tune steps of tests/golden/mono_freeze_tiny.npz."""
    import numpy as np
    import torch
    from collections import OrderedDict
    pre = OrderedDict((k[len("pre."):], torch.from_numpy(z[k].copy())) for k in z.files if k.startswith("pre."))
    pre_path = tmp_path / "snapshot.step.100"
    torch.save(pre, pre_path)
    s = make_solver_fn(pre_path)
    sd0 = {n: t.detach().cpu().clone() for n, t in s.asr_model.state_dict().items()}
    init = load_weights(tiny_cfg())
    # filter_model: pretrain_module entries come from the snapshot, the rest keep their initial values
    for n, t in sd0.items():
        if n == "pos_encoder.pe":
            continue
        src = pre[n] if n.split(".")[0] in ("feat_extractor", "vgg2enc", "encoder") else init[n]
        assert torch.equal(t, src), n
    assert all(not p.requires_grad for p in s.asr_model.encoder.parameters())
    assert all(p.requires_grad for p in s.asr_model.decoder.parameters())
    for step in range(int(z["n_steps"])):
        info = s.mono_step(step, *load_batch(z, f"s{step}."))
        ref = float(z[f"s{step}.loss"])
        assert abs(info["loss"] - ref) <= loss_rtol * abs(ref), (step, info["loss"], ref)
        assert abs(s.asr_opt.lr - float(z[f"s{step}.lr"])) < 1e-15
        for n, t in s.asr_model.state_dict().items():
            if n in ("pos_encoder.pe", "pre_embed.weight"):
                continue
            if n.split(".")[0] == "encoder":                     # frozen: bit-identical to the loaded snapshot
                assert torch.equal(t.detach().cpu(), sd0[n]), n
            else:
                check_adam_weights(z, f"s{step}.w.", [f"s{i}.g." for i in range(step + 1)], n, t.detach().cpu(),
                                   float(z[f"s{step}.lr"]), tol_lr=tol_lr)
    return s


def set_model_from_tiny_init(s):
    """set_model() with the golden's initial weights in place BEFORE load_model() overlays the pretrained modules
    (the reference initialises in MyTransformer.__init__, then TransformerTrainer.set_model calls load_model)."""
    from metaasr_crossaccent_b200 import interfaces as I
    orig = I.MonoMixin.load_model

    def load_with_init(self):
        self.asr_model.load_state_dict(load_weights(tiny_cfg()))
        orig(self)
    I.MonoMixin.load_model = load_with_init
    try:
        s.set_model()
    finally:
        I.MonoMixin.load_model = orig
    return s


def make_synth_accent_dir(root, seed, n_train=160, n_dev=24, idim=83):
    """Synthetic accent directory in the reference's on-disk format: <root>/{train,dev}/
    feat.dat (npy format), ilens.npy, label.npy, olens.npy.  Input lengths cluster on a few values so that the
    1-frame buckets of the train loader hold several utterances; feat[first frame of utterance i, 0] = i marks
    the utterance so that a batch reveals the dataset indices it was built from."""
    import numpy as np
    from pathlib import Path
    rng = np.random.default_rng(seed)
    for split, n in (("train", n_train), ("dev", n_dev)):
        d = Path(root, split)
        d.mkdir(parents=True, exist_ok=True)
        ilens = rng.choice(np.array([37, 40, 41, 44, 52, 53, 60, 75]), size=n).astype(np.int64)
        olens = rng.integers(1, 9, size=n).astype(np.int64)
        feat = rng.standard_normal((int(ilens.sum()), idim)).astype(np.float32)
        ptr = np.concatenate([[0], np.cumsum(ilens)])
        feat[ptr[:-1], 0] = np.arange(n, dtype=np.float32)
        label = rng.integers(1, 366, size=int(olens.sum())).astype(np.int64)
        with open(d / "feat.dat", "wb") as f:
            np.save(f, feat)
        np.save(d / "ilens.npy", ilens)
        np.save(d / "label.npy", label)
        np.save(d / "olens.npy", olens)
    return Path(root)


# ------------------------------------------------------------------ hkust-size cases (BASELINE shape B=32, T=512, L=32)
HKUST_SEED_W, HKUST_K, HKUST_WARMUP = 7, 0.2, 4


def hkust_profile_batch(seed, profile, B=32, T=512, L=32, idim=83):
    """Synthetic batch of SURVEY 8(d), regenerable from its seed on any box with this torch build:
    'eq'  -- what the reference's 1-frame-bucket train loader yields: all ilen = T, all label lengths = L;
    'rag' -- dev-loader-like: ilens = linspace(T -> T/2, B) sorted descending, L_b = ilen_b // 16, zeros beyond ilen."""
    g = torch.Generator().manual_seed(seed)
    if profile == "eq":
        ilens = torch.full((B,), T, dtype=torch.int64)
        lens = [L] * B
    else:
        ilens = torch.linspace(T, T // 2, B).round().to(torch.int64)
        lens = [int(t) // 16 for t in ilens.tolist()]
    x = torch.zeros(B, T, idim)
    for b, t in enumerate(ilens.tolist()):
        x[b, :t] = torch.randn(t, idim, generator=g)
    ys = [torch.randint(1, 366, (l,), generator=g, dtype=torch.int64) for l in lens]
    return x, ilens, ys, torch.tensor(lens, dtype=torch.int64)


def clone_batch(b):
    x, ilens, ys, olens = b
    return x.clone(), ilens.clone(), [y.clone() for y in ys], olens.clone()
