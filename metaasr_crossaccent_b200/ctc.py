"""Drop-in for the CTC loss of the reference's BLSTM trainer (src/blstm_trainer.py:22,65-70):

    self.ctc_loss = nn.CTCLoss(blank=0, reduction='mean', zero_infinity=True)
    loss = self.ctc_loss(pred.transpose(0,1).contiguous(), y_true, enc_lens, olens)   # pred = log_softmax

`B200CTCLoss` keeps nn.CTCLoss's call signature (log_probs [T,B,C], targets 1-D concat or [B,S],
input_lengths, target_lengths) and plugs into autograd through one fused forward+backward kernel
(masr_ctc_fwd_bwd): the gradient w.r.t. log_probs uses ATen's formula exp(lp) - posterior, so
back-propagating it through the caller's log_softmax gives the same logits gradient as the
reference.  `ctc_from_logits` is the fully fused variant (log-softmax inside the kernel).
"""
from __future__ import annotations

import torch

from . import _lib


def _prep(targets, target_lengths, device):
    tl = target_lengths.to("cpu", torch.int64)
    if targets.dim() == 2:                       # padded [B, S] form of nn.CTCLoss
        targets = torch.cat([targets[b, :int(tl[b])] for b in range(targets.shape[0])])
    offs = torch.zeros_like(tl)
    offs[1:] = torch.cumsum(tl, 0)[:-1]
    return (targets.to(device, torch.int64).contiguous(), offs.to(device), tl.to(device),
            int(tl.max()) if tl.numel() else 0)


def ctc_fwd_bwd(acts, targets, input_lengths, target_lengths, blank=0, zero_infinity=True,
                is_logprob=False, want_grad=True, grad_scale=1.0):
    """Raw call: returns (loss [1], nll [B], grad [T,B,C] or None), all on the device."""
    assert acts.is_cuda and acts.dtype == torch.float32 and acts.is_contiguous()
    lib = _lib.init(acts.device.index if acts.device.index is not None else torch.cuda.current_device())
    T, B, C = acts.shape
    tg, offs, tl, lmax = _prep(targets, target_lengths, acts.device)
    il = input_lengths.to(acts.device, torch.int64)
    nll = torch.empty(B, dtype=torch.float32, device=acts.device)
    loss = torch.empty(1, dtype=torch.float32, device=acts.device)
    grad = torch.empty_like(acts) if want_grad else None
    ws_bytes = lib.masr_ctc_workspace_bytes(T, B, C, lmax)
    ws = torch.empty(ws_bytes // 4 + 1, dtype=torch.float32, device=acts.device) if ws_bytes else None
    rc = lib.masr_ctc_fwd_bwd(acts.data_ptr(), T, B, C, int(is_logprob), tg.data_ptr(), offs.data_ptr(),
                              il.data_ptr(), tl.data_ptr(), lmax, blank, int(zero_infinity), float(grad_scale),
                              nll.data_ptr(), loss.data_ptr(), grad.data_ptr() if want_grad else None,
                              ws.data_ptr() if ws is not None else None, ws_bytes,
                              torch.cuda.current_stream(acts.device).cuda_stream)
    _lib.check(rc, "masr_ctc_fwd_bwd")
    return loss, nll, grad


class _CTCFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, acts, targets, input_lengths, target_lengths, blank, zero_infinity, is_logprob):
        loss, nll, grad = ctc_fwd_bwd(acts.contiguous(), targets, input_lengths, target_lengths, blank,
                                      zero_infinity, is_logprob, want_grad=True)
        ctx.save_for_backward(grad)
        return loss.squeeze(0)

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return grad * g, None, None, None, None, None, None


class B200CTCLoss(torch.nn.Module):
    """nn.CTCLoss(blank, reduction='mean', zero_infinity) replacement (reduction='mean' only: the
    reference's one configuration)."""

    def __init__(self, blank=0, reduction='mean', zero_infinity=True):
        super().__init__()
        if reduction != 'mean':
            raise NotImplementedError("B200CTCLoss implements reduction='mean' (src/blstm_trainer.py:22)")
        self.blank, self.zero_infinity = blank, zero_infinity

    def forward(self, log_probs, targets, input_lengths, target_lengths):
        return _CTCFunction.apply(log_probs, targets, input_lengths, target_lengths, self.blank,
                                  self.zero_infinity, True)


def ctc_from_logits(logits, targets, input_lengths, target_lengths, blank=0, zero_infinity=True):
    """Fused log_softmax + CTC (autograd-aware): logits [T, B, C] un-normalised."""
    return _CTCFunction.apply(logits, targets, input_lengths, target_lengths, blank, zero_infinity, False)
