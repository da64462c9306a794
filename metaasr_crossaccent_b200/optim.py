"""Optimizers of the hot path on flat arenas.

  noam schedule  : reference src/model/transformer_pytorch/optimizer.py:23-28
                   lr(n) = k * d_model^-0.5 * min(n^-0.5, n * warmup^-1.5)
  Adam           : torch.optim.Adam(betas=(0.9, 0.98), eps=1e-9), the only meta / multi-task optimizer the
                   reference builds (fo_meta_interface.py:103-109, transformer_torch_trainer.py:28-35)
  nesterov SGD   : torch.optim.SGD(lr=inner_lr, momentum, nesterov), re-created per task
                   (fo_meta_interface.py:228-236) -> here: one momentum arena re-used, `first_step` flag
"""
from __future__ import annotations

import torch


def noam_lr(step_num: int, k: float, d_model: int, warmup_steps: int) -> float:
    return k * (d_model ** -0.5) * min(step_num ** -0.5, step_num * warmup_steps ** -1.5)


class TransformerOptimizer:
    """noam learning-rate wrapper around any torch optimizer (API of the reference's wrapper:
    zero_grad / step / lr / step_num / state_dict / load_state_dict / set_k)."""

    def __init__(self, optimizer, k, d_model, warmup_steps=25000):
        self.optimizer, self.k, self.d_model, self.warmup_steps = optimizer, k, d_model, warmup_steps
        self.init_lr = d_model ** (-0.5)
        self.step_num, self.lr = 0, d_model ** (-0.5)

    def zero_grad(self):
        self.optimizer.zero_grad()

    def step(self):
        self.step_num += 1
        self.lr = noam_lr(self.step_num, self.k, self.d_model, self.warmup_steps)
        for group in self.optimizer.param_groups:
            group['lr'] = self.lr
        self.optimizer.step()

    def state_dict(self):
        return self.optimizer.state_dict()

    def load_state_dict(self, sd):
        self.optimizer.load_state_dict(sd)

    def set_k(self, k):
        self.k = k


class FlatAdamState:
    """Adam moments for one flat parameter arena, updated by the fused masr_mt_adam kernel."""

    def __init__(self, backend, params_flat, betas=(0.9, 0.98), eps=1e-9):
        self.be, self.p = backend, params_flat
        self.m = torch.zeros_like(params_flat)
        self.v = torch.zeros_like(params_flat)
        self.b1, self.b2 = betas
        self.eps, self.t = eps, 0

    def step(self, upd_flat, count, lr, skip_if_nan=None, clip_sumsq=None, max_norm=0.0, advance=True):
        if advance:
            self.t += 1
        bc1 = 1.0 - self.b1 ** self.t
        bc2 = 1.0 - self.b2 ** self.t
        self.be.mt_adam(self.p, self.m, self.v, upd_flat, count, lr, self.b1, self.b2, self.eps, bc1, bc2,
                        skip_if_nan, clip_sumsq, max_norm)


class FlatNoamAdam:
    """noam-Adam over the engine's parameter arena: the `asr_opt` of multi-task / mono training
    (multi_interface.py:108-114).  step() consumes engine.grads; with `gnorm_sumsq` the
    clip_grad_norm_(GRAD_CLIP) scaling and the NaN guard are fused into the same kernel."""

    def __init__(self, engine, k, d_model, warmup_steps):
        self.engine, self.k, self.d_model, self.warmup_steps = engine, k, d_model, warmup_steps
        self.state = FlatAdamState(engine.be, engine.params)
        self.step_num, self.lr = 0, d_model ** (-0.5)
        self.param_groups = [{'lr': self.lr}]

    def zero_grad(self):
        pass                      # the gradient arena is zeroed by the engine at the start of every backward

    def step(self, gnorm_sumsq=None, max_norm=0.0):
        self.step_num += 1
        self.lr = noam_lr(self.step_num, self.k, self.d_model, self.warmup_steps)
        self.param_groups[0]['lr'] = self.lr
        self.state.step(self.engine.grads, 1.0, self.lr, skip_if_nan=gnorm_sumsq, clip_sumsq=gnorm_sumsq,
                        max_norm=max_norm)
        self.engine.weights_dirty = True

    def state_dict(self):
        return {"m": self.state.m, "v": self.state.v, "t": self.state.t, "step_num": self.step_num}

    def load_state_dict(self, sd):
        self.state.m.copy_(sd["m"]); self.state.v.copy_(sd["v"])
        self.state.t, self.step_num = sd["t"], sd["step_num"]


class FlatInnerSGD:
    """Inner-loop optimizer of run_task: clip_grad_norm_ + nesterov SGD fused (masr_mt_clip_sgd)."""

    def __init__(self, engine, lr, momentum, nesterov):
        self.engine, self.lr, self.momentum, self.nesterov = engine, lr, momentum, nesterov
        self.buf = torch.zeros_like(engine.params)
        self.first = True

    def reset(self):
        self.first = True         # fresh SGD per task: momentum never carries across tasks

    def zero_grad(self):
        pass

    def step(self, gnorm_sumsq, max_norm, last=False):
        """last: no further step of this task follows and the caller reads neither .grad nor the momentum again (the
        lock-step meta-step scheduler): the kernel skips those write-backs.  On the CUDA path the same pass writes the
        engine's bf16 shadow of the new weights."""
        e = self.engine
        if hasattr(e.be, "mt_clip_sgd_ex"):
            n = e.layout.total
            wrote = e.be.mt_clip_sgd_ex(e.params[:n], e.grads[:n], self.buf[:n], gnorm_sumsq, max_norm, self.lr, self.momentum,
                                        self.nesterov, self.first, None if e.shadow is None else e.shadow[:n], last)
            self.first = False
            if wrote:
                e.mark_shadow_fresh()
            else:
                e.weights_dirty = True
            return
        e.be.mt_clip_sgd(e.params, e.grads, self.buf, gnorm_sumsq, max_norm, self.lr, self.momentum,
                         self.nesterov, self.first)
        self.first = False
        e.weights_dirty = True
