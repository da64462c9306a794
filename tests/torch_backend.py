"""TEST DOUBLE (test infrastructure, never shipped): the kernel-backend interface of
metaasr_crossaccent_b200.ops.CudaBackend re-stated with plain torch ops, so that the host-side
orchestration in engine.py / interfaces.py (which tensor feeds which kernel, accumulation flags,
buffer reuse, arena layout) can be checked against the oracle on a CPU-only box.  Each method
documents the contract of the CUDA kernel of the same name; the `-m gpu` tests check the kernels
themselves against this file's semantics.  Dropout must be 0 here (the device RNG cannot be
reproduced on the host)."""
from __future__ import annotations

import torch
import torch.nn.functional as F


class TorchBackend:
    name = "torch-test-double"

    def __init__(self, device="cpu", act_dtype=torch.float32):
        self.device = torch.device(device)
        self.act_dtype = act_dtype
        self.launches = 0

    # ---- GEMM family
    def linear_fwd(self, x, w, bias, y, relu=False, dropout=None):
        assert dropout is None or dropout[0] == 0.0
        o = x.float() @ w.float().t()
        if bias is not None:
            o = o + bias
        y.copy_(F.relu(o) if relu else o)

    def linear_dgrad(self, dy, w, dx, accumulate=False, relu_drop_mask=None, p=0.0, rowdot=None):
        if rowdot is not None:                     # the test double leaves D to attn_bwd
            self.linear_dgrad(dy, w, dx)
            return False
        o = dy.float() @ w.float()
        if relu_drop_mask is not None:
            assert not accumulate
            o = o * (relu_drop_mask.float() > 0) / (1.0 - p)
        dx.copy_(dx.float() + o if accumulate else o)

    def linear_wgrad(self, x, dy, dw, db):
        dw += dy.float().t() @ x.float()
        if db is not None:
            db += dy.float().sum(0)

    # ---- conv front end (NHWC activations; wp [Cout, tap*Cin + ci])
    @staticmethod
    def _unprep(wp, Cin):
        Cout = wp.shape[0]
        return wp.float().view(Cout, 3, 3, Cin).permute(0, 3, 1, 2).contiguous()     # [Cout,Cin,3,3]

    def conv1_fwd(self, x, w, bias, y):
        o = F.relu(F.conv2d(x.unsqueeze(1), w.view(-1, 1, 3, 3), bias, padding=1))     # [B,C,H,W]
        y.copy_(o.permute(0, 2, 3, 1))

    def conv1_wgrad(self, x, dy, dw, db):
        g = dy.float().permute(0, 3, 1, 2)
        gw = torch.nn.grad.conv2d_weight(x.unsqueeze(1), (dw.shape[0], 1, 3, 3), g, padding=1)
        dw += gw.view(dw.shape)
        db += g.sum((0, 2, 3))

    def conv_w_prep(self, w, wp):
        wp.copy_(w.permute(0, 2, 3, 1).reshape(w.shape[0], -1))

    def conv_w_unprep_add(self, dwp, dw):
        dw += self._unprep(dwp, dw.shape[1])

    def conv3x3_fwd(self, x, wp, bias, y):
        w = self._unprep(wp, x.shape[3])
        o = F.relu(F.conv2d(x.float().permute(0, 3, 1, 2), w, bias, padding=1))
        y.copy_(o.permute(0, 2, 3, 1))

    def conv3x3_dgrad(self, dy, wp, dx, relu_src=None):
        w = self._unprep(wp, dx.shape[3])
        g = F.conv_transpose2d(dy.float().permute(0, 3, 1, 2), w, padding=1).permute(0, 2, 3, 1)
        if relu_src is not None:
            g = g * (relu_src.float() > 0)
        dx.copy_(g)

    def conv3x3_wgrad(self, x, dy, dwp, db):
        g = dy.float().permute(0, 3, 1, 2)
        Cout, Cin = dy.shape[3], x.shape[3]
        gw = torch.nn.grad.conv2d_weight(x.float().permute(0, 3, 1, 2), (Cout, Cin, 3, 3), g, padding=1)
        dwp += gw.permute(0, 2, 3, 1).reshape(Cout, -1)
        db += g.sum((0, 2, 3))

    def maxpool_fwd(self, x, y):
        y.copy_(F.max_pool2d(x.float().permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1))

    def maxpool_bwd(self, x, dy, dx, relu_mask=True):
        xx = x.float().permute(0, 3, 1, 2).detach().requires_grad_(True)
        o = F.max_pool2d(xx, 2, 2)
        (g,) = torch.autograd.grad(o, xx, dy.float().permute(0, 3, 1, 2))
        g = g.permute(0, 2, 3, 1)
        if relu_mask:
            g = g * (x.float() > 0)
        dx.copy_(g)

    def relu_bwd(self, y, dx):
        dx.mul_((y.float() > 0).to(dx.dtype))

    # ---- attention (masks from lengths)
    @staticmethod
    def _mask(B, Lq, Lk, klens, causal, device):
        m = torch.zeros(B, 1, Lq, Lk, device=device)
        if klens is not None:
            pad = torch.arange(Lk, device=device)[None, :] >= klens[:, None]
            m = m.masked_fill(pad[:, None, None, :], float("-inf"))
        if causal:
            m = m + torch.triu(torch.full((Lq, Lk), float("-inf"), device=device), diagonal=1)
        return m

    def _attn(self, q, k, v, B, H, Lq, Lk, klens, causal):
        d = q.shape[1]
        hd = d // H
        qq = q.float().reshape(B, Lq, H, hd).transpose(1, 2)
        kk = k.float().reshape(B, Lk, H, hd).transpose(1, 2)
        vv = v.float().reshape(B, Lk, H, hd).transpose(1, 2)
        s = qq @ kk.transpose(-1, -2) / (hd ** 0.5) + self._mask(B, Lq, Lk, klens, causal, q.device)
        return qq, kk, vv, s

    def attn_fwd(self, q, k, v, out, lse, B, H, Lq, Lk, klens, causal, p=0.0, seed=0, site=0, kv_rows=None):
        assert p == 0.0
        if kv_rows is not None and kv_rows != Lk:        # decode cache of fixed capacity: the first Lk rows per utterance
            k = k.reshape(B, kv_rows, -1)[:, :Lk].reshape(B * Lk, -1)
            v = v.reshape(B, kv_rows, -1)[:, :Lk].reshape(B * Lk, -1)
        qq, kk, vv, s = self._attn(q, k, v, B, H, Lq, Lk, klens, causal)
        lse.copy_(torch.logsumexp(s, -1).reshape(-1))
        o = torch.softmax(s, -1) @ vv
        out.copy_(o.transpose(1, 2).reshape(B * Lq, -1))

    def attn_bwd(self, q, k, v, out, dout, lse, dsum, dq, dk, dv, B, H, Lq, Lk, klens, causal, p=0.0, seed=0, site=0,
                 dsum_ready=False):
        assert p == 0.0
        qd, kd, vd = (t.float().detach().clone().requires_grad_(True) for t in (q, k, v))
        qq, kk, vv, s = self._attn(qd, kd, vd, B, H, Lq, Lk, klens, causal)
        o = (torch.softmax(s, -1) @ vv).transpose(1, 2).reshape(B * Lq, -1)
        gq, gk, gv = torch.autograd.grad(o, (qd, kd, vd), dout.float())
        dq.copy_(gq); dk.copy_(gk); dv.copy_(gv)

    # ---- fused elementwise
    def add_layernorm_fwd(self, x, res, gamma, beta, y, mean, rstd, p=0.0, seed=0, site=0, eps=1e-5):
        assert p == 0.0
        s = x.float() + (res.float() if res is not None else 0)
        x.copy_(s)
        s = x.float()
        mu = s.mean(-1)
        var = s.var(-1, unbiased=False)
        r = torch.rsqrt(var + eps)
        mean.copy_(mu); rstd.copy_(r)
        y.copy_((s - mu[:, None]) * r[:, None] * gamma + beta)

    def add_layernorm_bwd(self, dy, s, mean, rstd, gamma, ds, ds_accum, dx, dgamma, dbeta, p=0.0, seed=0, site=0):
        assert p == 0.0
        xh = (s.float() - mean[:, None]) * rstd[:, None]
        g = dy.float() * gamma
        v = rstd[:, None] * (g - g.mean(-1, keepdim=True) - xh * (g * xh).mean(-1, keepdim=True))
        dgamma += (dy.float() * xh).sum(0)
        dbeta += dy.float().sum(0)
        if dx is not None:
            dx.copy_(v)
        ds.copy_(ds.float() + v if ds_accum else v)

    def add_pe_dropout(self, x, pe, L, p=0.0, seed=0, site=0):
        assert p == 0.0
        rows, d = x.shape
        x.copy_((x.float().view(-1, L, d) + pe[:L]).view(rows, d))

    def embed_pe_fwd(self, ids, E, pe, out, L, p=0.0, seed=0, site=0):
        assert p == 0.0
        rows, d = out.shape
        out.copy_((E[ids].view(-1, L, d) + pe[:L]).view(rows, d))

    def embed_bwd(self, ids, dout, dE, L, p=0.0, seed=0, site=0):
        assert p == 0.0
        dE.index_add_(0, ids, dout.float())

    def dropout(self, x, p, seed, site):
        assert p == 0.0

    def colsum_add(self, x, out):
        out += x.float().sum(0)

    def cast(self, src, dst):
        dst.copy_(src)

    def permute_cf(self, src, dst, Cc, Fq, inverse_add=False):
        rows = src.shape[0]
        if not inverse_add:
            dst.copy_(src.view(rows, Cc, Fq).transpose(1, 2).reshape(rows, -1))
        else:
            dst += src.view(rows, Fq, Cc).transpose(1, 2).reshape(rows, -1)

    def ls_ce(self, logits, gold, eps, inv_n, stats, argmax, dlogits, inv_n_dev=None):
        if inv_n_dev is not None:
            inv_n = float(inv_n_dev[0])
        N, Cc = logits.shape
        keep = gold >= 0
        lz = logits.detach().clone().requires_grad_(True)
        gs = keep.long() * gold
        one_hot = torch.zeros_like(lz).scatter(1, gs.view(-1, 1), 1)
        q = one_hot * (1 - eps) + (1 - one_hot) * eps / Cc
        rows = -(q * F.log_softmax(lz, -1)).sum(1)
        tot = rows.masked_select(keep).sum()
        am = logits.max(1)[1]
        stats[0] += tot.double().item()
        stats[1] += float((am.eq(gold) & keep).sum())
        stats[2] += float(keep.sum())
        if argmax is not None:
            argmax.copy_(am)
        if dlogits is not None:
            (g,) = torch.autograd.grad(tot * inv_n, lz)
            dlogits.copy_(g)

    def ctc_joint(self, logits, B, T, Cc, targets, offs, in_lens, tgt_lens, lmax, w, nll, loss, grad):
        """Torch restatement of the joint objective's CTC branch (espnet-style): F.ctc_loss(log_softmax(logits)) with
        blank 0, reduction 'mean', zero_infinity on the batch-first logits."""
        lg = logits.detach().double().view(B, T, Cc).requires_grad_(True)
        lp = torch.log_softmax(lg, -1).transpose(0, 1)
        tl = tgt_lens.cpu().long()
        tg = torch.cat([targets[int(offs[b]):int(offs[b]) + int(tl[b])] for b in range(B)]).long().cpu()
        l = torch.nn.functional.ctc_loss(lp, tg, in_lens.cpu().long(), tl, blank=0, reduction='mean', zero_infinity=True)
        loss.copy_(l.detach().float().view(1))
        if grad is not None:
            (g,) = torch.autograd.grad(l, lg)
            grad.copy_((w * g).float().view(B * T, Cc))

    def cast_pad2d(self, src, dst, cols):
        dst.zero_()
        dst[:, :cols].copy_(src[:, :cols])

    def loss_mix(self, stats, ctc_loss, w):
        stats[3] = float(ctc_loss)
        stats[4] = w

    def zero_(self, t):
        t.zero_()

    # ---- flat arena ops (kernel 4)
    @staticmethod
    def _coef(sumsq, max_norm):
        total = float(sumsq.sqrt())
        return min(1.0, max_norm / (total + 1e-6)) if total == total else 1.0

    def mt_sumsq(self, g, out, zero_first=True):
        if zero_first:
            out.zero_()
        out += g.double().pow(2).sum()

    def mt_clip_sgd(self, p, g, buf, sumsq, max_norm, lr, momentum, nesterov, first_step):
        if bool(torch.isnan(sumsq).any()):
            return
        g.mul_(self._coef(sumsq[0], max_norm))
        d = g
        if momentum != 0:
            if first_step:
                buf.copy_(g)
            else:
                buf.mul_(momentum).add_(g)
            d = g.add(buf, alpha=momentum) if nesterov else buf
        p.add_(d, alpha=-lr)

    def mt_clip(self, g, sumsq, max_norm):
        g.mul_(self._coef(sumsq[0], max_norm))

    def mt_accumulate(self, upd, g, sumsq=None, max_norm=0.0):
        upd.add_(g, alpha=self._coef(sumsq[0], max_norm) if sumsq is not None else 1.0)

    def mt_reptile_delta(self, upd, theta, phi):
        upd.add_(theta - phi)

    def mt_adam(self, p, m, v, upd, count, lr, beta1, beta2, eps, bc1, bc2, skip_if_nan=None, clip_sumsq=None, max_norm=0.0):
        if skip_if_nan is not None and bool(torch.isnan(skip_if_nan).any()):
            return
        g = upd / count
        if clip_sumsq is not None:
            g = g * self._coef(clip_sumsq[0], max_norm)
        m.lerp_(g, 1 - beta1)
        v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
        denom = (v.sqrt() / (bc2 ** 0.5)).add_(eps)
        p.addcdiv_(m, denom, value=-(lr / bc1))

    def mt_axpy(self, y, x, a):
        y.add_(x, alpha=a)

    def copy_(self, dst, src):
        dst.copy_(src)
