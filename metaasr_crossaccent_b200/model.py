"""B200Transformer: the nn.Module face of the engine.

Keeps the reference module's contract (src/model/transformer_pytorch/mono_transformer_torch.py):
same 114 state-dict keys / shapes / fp32 dtype, `parameters()` yields 112 unique
tensors with the tied matrix once, `forward(xs_pad, ilens, ys, olens) -> (logit, ys_out_pad)`,
`recog(xs_pad, ilens) -> ids[L, B]`, `sos_id`, `eos_id`.  Every Parameter is a VIEW into the
engine's flat fp32 arena and its `.grad` a view into the gradient arena, so reference-style code
(`clip_grad_norm_(model.parameters())`, `torch.optim.SGD(model.parameters())`,
`load_state_dict(_original)`) keeps working, while the fused interfaces operate on the arenas.
"""
from __future__ import annotations

import torch
from torch import nn

from .engine import NetConfig, TransformerEngine


class B200Transformer(nn.Module):
    def __init__(self, id2char, model_para, backend, device, label_smoothing=0.2, seed=531, init=True):
        super().__init__()
        self.cfg = NetConfig.from_yaml(model_para, odim=len(id2char))
        self.idim, self.odim = self.cfg.idim, self.cfg.odim
        self.sos_id, self.eos_id = self.cfg.sos_id, self.cfg.eos_id
        self.d_model, self.nhead = self.cfg.d_model, self.cfg.nheads
        self.engine = TransformerEngine(self.cfg, backend, device, label_smoothing, seed)
        eng = self.engine
        made = {}
        for name in eng.layout.shapes:
            path = name.split(".")
            mod = self
            for part in path[:-1]:
                if part not in mod._modules:
                    mod.add_module(part, nn.Module())
                mod = mod._modules[part]
            if name == "pos_encoder.pe":
                mod.register_buffer("pe", eng.pe)
                continue
            src = "char_trans.weight" if (self.cfg.tie and name == "pre_embed.weight") else name
            if src not in made:
                p = nn.Parameter(eng.P[src], requires_grad=True)
                p.grad = eng.G[src]
                made[src] = p
            mod.register_parameter(path[-1], made[src])
        if init:
            self.init_parameters(seed)

    @property
    def device(self):
        return self.engine.device

    def init_parameters(self, seed=None):
        """Reference init (mono_transformer_torch.py:106-109): xavier_uniform_ on dim > 1; vectors
        keep torch's module defaults (LN weight 1, attention/LN biases 0, conv/linear biases
        U(+-1/sqrt(fan_in))).  Drawn on the host for device independence."""
        import math
        g = torch.Generator().manual_seed(531 if seed is None else int(seed))
        with torch.no_grad():
            for name, p in self.named_parameters():
                shape = tuple(p.shape)
                if p.dim() > 1:
                    rf = shape[2] * shape[3] if p.dim() == 4 else 1
                    a = math.sqrt(6.0 / (shape[1] * rf + shape[0] * rf))
                    v = (torch.rand(shape, generator=g) * 2 - 1) * a
                elif "norm" in name and name.endswith("weight"):
                    v = torch.ones(shape)
                elif "norm" in name or "in_proj_bias" in name or "out_proj.bias" in name:
                    v = torch.zeros(shape)
                else:
                    w = dict(self.named_parameters())[name[:-4] + "weight"]
                    fan_in = w.shape[1] * (w.shape[2] * w.shape[3] if w.dim() == 4 else 1)
                    v = (torch.rand(shape, generator=g) * 2 - 1) / math.sqrt(fan_in)
                p.copy_(v.to(p.device))
        self.engine.weights_dirty = True

    def load_state_dict(self, state_dict, strict=True):
        out = super().load_state_dict(state_dict, strict=strict)
        self.engine.weights_dirty = True
        return out

    def attach_grads(self):
        """Re-point every Parameter's .grad at its gradient-arena view (an external
        optimizer.zero_grad(set_to_none=True) detaches them)."""
        eng = self.engine
        first = next(iter(self.parameters()))
        if first.grad is not None and first.grad.data_ptr() == eng.grads.data_ptr():
            return
        for name, p in self.named_parameters():
            p.grad = eng.G[name]

    def train(self, mode=True):
        super().train(mode)
        self.engine.training = bool(mode)
        return self

    # -- MyTransformer.recog (:143-176): greedy decoding.  Like the reference, the decoder is re-run on the growing
    # prefix for max(enc_lens) steps and EVERY position is re-decided each time; unlike the reference the conv
    # front end and the encoder run once (their output is cached and handed to the later steps).
    @torch.no_grad()
    def recog(self, xs_pad, ilens, kv_cache=True):
        """Greedy ids [L, B].  kv_cache=True (default): engine.greedy_decode -- one decoder row per step against cached
        keys / values; kv_cache=False: the reference's own schedule (decoder re-run on the growing prefix), kept as the
        cross-check of the cached path."""
        eng = self.engine
        if kv_cache:
            n_steps = int(torch.floor(ilens.to(dtype=torch.float32) / 4).max())
            if n_steps <= 0:
                return torch.zeros((0, xs_pad.shape[0]), dtype=torch.int64)
            return eng.greedy_decode(xs_pad, ilens, n_steps).cpu()
        eng.weights_dirty = True
        was_training, eng.training = eng.training, False
        B = xs_pad.shape[0]
        out = torch.zeros((B, 0), dtype=torch.int64)
        n_steps = int(torch.floor(ilens.to(dtype=torch.float32) / 4).max())
        mem = None
        for _ in range(n_steps):
            hb = eng.prepare_batch(xs_pad, ilens, [out[b] for b in range(B)], None)
            ws = eng.forward(eng.to_device(hb), want_grad=False, mem=mem)
            if mem is None:
                mem = ws["mem"].clone()
            out = ws["argmax"].view(B, hb["L1"]).cpu()
        eng.training = was_training
        return out.t().contiguous()          # [L, B] like the reference

    # -- MyTransformer.forward (:178-208); inference-style (no autograd graph): logits on device
    @torch.no_grad()
    def forward(self, xs_pad, ilens, ys, olens):
        eng = self.engine
        eng.weights_dirty = True
        hb = eng.prepare_batch(xs_pad, ilens, ys, olens)
        db = eng.to_device(hb)
        ws = eng.forward(db, want_grad=False)
        logit = ws["logits"].view(hb["B"], hb["L1"], self.odim)
        return logit, db["ys_out"]
