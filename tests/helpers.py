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
