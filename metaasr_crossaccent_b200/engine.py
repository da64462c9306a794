"""Forward + backward of the transformer_pytorch ASR model as an explicit schedule of C-ABI kernel
calls over flat parameter / gradient arenas (no autograd, no torch ops on the path).

Mirrors MyTransformer.forward (reference src/model/transformer_pytorch/mono_transformer_torch.py:
113-141,178-208) and the loss of TransformerTrainer.run_batch (src/transformer_torch_trainer.py:
59-92), including the reference's quirks: padded frames are not re-masked in the
conv stack, floor-mode pooling, enc_lens = floor(ilens/4), no target key-padding in decoder
self-attention, no sqrt(d) embedding scale, label smoothing over n_class, tied output projection.

Layout decisions (B200-first):
  * every trainable tensor is a view into ONE flat fp32 arena (state-dict order, tied matrix once),
    gradients into a second arena of the same layout -> clip / SGD / FOMAML accumulate / Adam /
    all-reduce are single streaming passes (ops.mt_*);
  * activations are batch-first [B*T, d]; conv tensors NHWC so that (a) the implicit-GEMM K dim
    (tap, Cin) is contiguous and (b) the pooled conv4 output IS the [B*T', 2560] vgg2enc input
    (the (c,f)->(f,c) flattening difference is folded into a column permutation of vgg2enc.weight);
  * masks are never materialised: kernels take lengths.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import torch

IGNORE_ID = -1
LN_EPS = 1e-5
ARENA_ALIGN = 64          # elements (256 B): keeps every tensor 16 B aligned for 128-bit / TMA access


class NetConfig:
    """asr_model block of the reference YAMLs (config/transformer/pretrain/fometa-hkust.yaml:13-26)."""

    def __init__(self, idim=83, d_model=512, nheads=8, d_inner=2048, enc_layers=2, dec_layers=4,
                 odim=367, dropout=0.1, pos_dropout=0.1, tie=True, ctc_weight=0.0):
        self.idim, self.d_model, self.nheads, self.d_inner = idim, d_model, nheads, d_inner
        self.enc_layers, self.dec_layers, self.odim = enc_layers, dec_layers, odim
        self.dropout, self.pos_dropout, self.tie = float(dropout), float(pos_dropout), bool(tie)
        self.sos_id, self.eos_id = 0, odim - 1
        # joint CTC / attention objective (north_star kernel 3; an EXTENSION -- the reference only carries dead config for
        # it, config/transformer/mono-test.yaml:44-50): total = (1-w) * LS-CE + w * CTC(ctc_lo(encoder memory), ys),
        # blank = index 0.  w = 0 (default): no CTC head, state dict and numerics are exactly the reference's.
        self.ctc_weight = float(ctc_weight)
        assert 0.0 <= self.ctc_weight < 1.0
        self.vgg_ch = 128
        self.f4 = idim // 4
        self.vgg_o_dim = self.vgg_ch * self.f4
        assert d_model % nheads == 0 and d_model // nheads <= 64

    @staticmethod
    def from_yaml(am: dict, odim: int):
        return NetConfig(am["idim"], am["d_model"], am["nheads"], am["d_inner"], am["encoder"]["nlayers"],
                         am["decoder"]["nlayers"], odim, am.get("dropout", 0.0), am.get("pos_dropout", 0.0),
                         am.get("tgt_share_weight", 1) != 0, am.get("ctc_weight", 0.0))


def param_shapes(cfg: NetConfig) -> "OrderedDict[str, tuple]":
    """state_dict key -> shape in the reference's registration order (114 entries for the hkust net)."""
    d, ff = cfg.d_model, cfg.d_inner
    s = OrderedDict()
    for i, (co, ci) in zip((0, 2, 5, 7), ((64, 1), (64, 64), (128, 64), (128, 128))):
        s[f"feat_extractor.{i}.weight"] = (co, ci, 3, 3)
        s[f"feat_extractor.{i}.bias"] = (co,)
    s["vgg2enc.weight"] = (d, cfg.vgg_o_dim)
    s["vgg2enc.bias"] = (d,)
    s["pos_encoder.pe"] = (3000, 1, d)
    s["char_trans.weight"] = (cfg.odim, d)
    s["char_trans.bias"] = (cfg.odim,)
    s["pre_embed.weight"] = (cfg.odim, d)

    def attn(p):
        s[p + ".in_proj_weight"] = (3 * d, d)
        s[p + ".in_proj_bias"] = (3 * d,)
        s[p + ".out_proj.weight"] = (d, d)
        s[p + ".out_proj.bias"] = (d,)

    def ffn_norms(p, n_norm):
        s[p + ".linear1.weight"] = (ff, d)
        s[p + ".linear1.bias"] = (ff,)
        s[p + ".linear2.weight"] = (d, ff)
        s[p + ".linear2.bias"] = (d,)
        for j in range(1, n_norm + 1):
            s[p + f".norm{j}.weight"] = (d,)
            s[p + f".norm{j}.bias"] = (d,)

    for l in range(cfg.enc_layers):
        attn(f"encoder.layers.{l}.self_attn")
        ffn_norms(f"encoder.layers.{l}", 2)
    s["encoder.norm.weight"] = (d,)
    s["encoder.norm.bias"] = (d,)
    for l in range(cfg.dec_layers):
        attn(f"decoder.layers.{l}.self_attn")
        attn(f"decoder.layers.{l}.multihead_attn")
        ffn_norms(f"decoder.layers.{l}", 3)
    s["decoder.norm.weight"] = (d,)
    s["decoder.norm.bias"] = (d,)
    if cfg.ctc_weight > 0.0:          # appended LAST: every reference tensor keeps its arena offset
        s["ctc_lo.weight"] = (cfg.odim, d)
        s["ctc_lo.bias"] = (cfg.odim,)
    return s


def positional_table(max_len, d_model):
    """PositionalEncoding buffer, computed exactly as mono_transformer_torch.py:21-27 (fp32, CPU)."""
    pe = torch.zeros(max_len, d_model)
    position = torch.arange(0, max_len, dtype=torch.float).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d_model, 2).float() * (-math.log(10000.0) / d_model))
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe.unsqueeze(0).transpose(0, 1).contiguous()      # [max_len, 1, d]


class ArenaLayout:
    """Offsets of the unique trainable tensors inside the flat arenas."""

    def __init__(self, cfg: NetConfig):
        self.shapes = param_shapes(cfg)
        self.offsets = OrderedDict()
        off = 0
        for name, shape in self.shapes.items():
            if name == "pos_encoder.pe" or (cfg.tie and name == "pre_embed.weight"):
                continue
            self.offsets[name] = off
            n = 1
            for v in shape:
                n *= v
            off += (n + ARENA_ALIGN - 1) // ARENA_ALIGN * ARENA_ALIGN
        self.total = off
        self.n_unique_elems = sum(int(torch.Size(self.shapes[n]).numel()) for n in self.offsets)

    def view(self, arena: torch.Tensor, name: str) -> torch.Tensor:
        shape = self.shapes[name]
        off = self.offsets[name]
        return arena[off:off + int(torch.Size(shape).numel())].view(shape)


class TransformerEngine:
    """Owns arenas + workspaces and runs forward / backward through a kernel backend."""

    def __init__(self, cfg: NetConfig, backend, device, label_smoothing=0.2, seed=531):
        self.cfg, self.be, self.device = cfg, backend, torch.device(device)
        self.eps_ls = float(label_smoothing)
        self.act_dtype = backend.act_dtype
        self.layout = ArenaLayout(cfg)
        n = self.layout.total
        self.params = torch.zeros(n, dtype=torch.float32, device=self.device)
        self.grads = torch.zeros(n, dtype=torch.float32, device=self.device)
        self.pe = positional_table(3000, cfg.d_model).to(self.device)               # [3000,1,d]
        self.pe2d = self.pe.view(3000, cfg.d_model)
        self.P = OrderedDict((nm, self.layout.view(self.params, nm)) for nm in self.layout.offsets)
        self.G = OrderedDict((nm, self.layout.view(self.grads, nm)) for nm in self.layout.offsets)
        if cfg.tie:
            self.P["pre_embed.weight"] = self.P["char_trans.weight"]
            self.G["pre_embed.weight"] = self.G["char_trans.weight"]
        # compute-dtype shadow of the parameter arena (bf16 mode) and re-laid-out conv / vgg2enc weights
        self.shadow = None
        if self.act_dtype != torch.float32:
            self.shadow = torch.zeros(n, dtype=self.act_dtype, device=self.device)
        self.W = OrderedDict()      # name -> compute-dtype weight used by GEMMs
        self._shadow_views = None
        self._shadow_fresh, self._shadow_version = False, -1
        self.wpt = {}               # conv idx -> [Cin, 9*Cout] compute dtype (dgrad operand), tensor-core path only
        self.wp = {}                # conv idx -> [Cout, 9*Cin] compute dtype
        self.dwp = {}               # conv idx -> fp32 gradient in the same layout
        for i, (co, ci) in zip((2, 5, 7), ((64, 64), (128, 64), (128, 128))):
            self.wp[i] = torch.zeros(co, 9 * ci, dtype=self.act_dtype, device=self.device)
            self.dwp[i] = torch.zeros(co, 9 * ci, dtype=torch.float32, device=self.device)
            if hasattr(backend, "conv_w_prep_t") and getattr(backend, "gemm_path", "") == "umma":
                self.wpt[i] = torch.zeros(ci, 9 * co, dtype=self.act_dtype, device=self.device)
        self.vgg2enc_p = torch.zeros(cfg.d_model, cfg.vgg_o_dim, dtype=self.act_dtype, device=self.device)
        self.d_vgg2enc_p = torch.zeros(cfg.d_model, cfg.vgg_o_dim, dtype=torch.float32, device=self.device)
        self.weights_dirty = True
        # [sum of row losses, n_correct, n_non_pad, CTC term, ctc_weight, -, -, -]
        self.stats = torch.zeros(8, dtype=torch.float64, device=self.device)
        # device-resident dropout seed offset of THIS engine (kernels add it to every dropout seed; a captured
        # CUDA graph bumps it at replay).  Per engine, so that engines running concurrently on different
        # streams cannot change each other's masks between a forward and its backward.
        self.seed_t = torch.zeros(1, dtype=torch.int64, device=self.device)
        self.step_seed = int(seed) * 1000003
        self._ws = {}
        self._sites = {}
        self.training = True
        self.use_graphs = False           # CUDA-graph replay of forward+backward per (B, T, L1) shape
        self._graphs = OrderedDict()
        self.max_graphs = 8
        # weight-gradient GEMMs (and bias sums) never feed the back-propagation chain: they run on a
        # second stream, concurrently with the dgrad chain (small launches fill only part of the 148 SMs)
        self.multi_stream = self.device.type == "cuda"
        # grouped launch of the weight-gradient GEMMs (backward_rest).  Pays when several task lanes share the GPU (one
        # ~1 500-CTA grid instead of 26 small launches per lane: 31.9 -> 31.6 ms at four lanes); on a single lane the small
        # launches hide under the latency-bound dgrad chain and deferring them costs 0.05 ms per batch, so the meta-step
        # scheduler switches it on only for multi-lane steps
        self.group_wgrads = False
        self._side = None
        self._side_busy = False

    # ------------------------------------------------------------------ parameters
    def load_state_dict(self, sd):
        """Copy a reference-layout state dict (114 keys) into the arena; with tied weights the
        pre_embed.weight entry is written last, as nn.Module.load_state_dict does."""
        for name in self.layout.shapes:
            if name == "pos_encoder.pe" or name not in sd:
                continue
            self.P[name].copy_(sd[name].to(self.device, torch.float32))
        self.weights_dirty = True

    def state_dict(self):
        """Reference-compatible state dict (same keys / shapes / fp32), tensors are arena views."""
        out = OrderedDict()
        for name in self.layout.shapes:
            out[name] = self.pe if name == "pos_encoder.pe" else self.P[name]
        return out

    def mark_shadow_fresh(self):
        """An optimizer kernel (masr_mt_clip_sgd_ex / masr_mt_copy_cast) just wrote the compute-dtype shadow together with
        the master arena: the next prep_weights skips its cast pass.  Consumed by that one prep; anything else that
        touches the parameters (load_state_dict, a foreign optimizer on the nn.Parameters) goes through the full prep."""
        self.weights_dirty = True
        self._shadow_fresh = True
        self._shadow_version = self.params._version      # a torch-side in-place write to the arena (or a view) bumps it

    def prep_weights(self):
        """Refresh the derived weight copies after the fp32 master arena changed."""
        if not self.weights_dirty:
            return
        be = self.be
        src = self.P
        fresh = self._shadow_fresh and self._shadow_version == self.params._version
        self._shadow_fresh = False
        if self.shadow is not None:
            src = self._shadow_views
            if src is None:
                src = self._shadow_views = OrderedDict((nm, self.layout.view(self.shadow, nm)) for nm in self.layout.offsets)
        self.W = src
        if hasattr(be, "prep_weights"):            # one launch: cast (unless fresh) + conv re-layouts + vgg2enc permutation
            jobs = [(self.P[f"feat_extractor.{i}.weight"], self.wp[i], self.wpt.get(i)) for i in (2, 5, 7)]
            be.prep_weights(self.params, None if (self.shadow is None or fresh) else self.shadow, jobs,
                            self.P["vgg2enc.weight"], self.vgg2enc_p, self.cfg.vgg_ch, self.cfg.f4, self.vgg2enc_p)
        else:
            if self.shadow is not None and not fresh:
                be.cast(self.params, self.shadow)
            for i in (2, 5, 7):
                be.conv_w_prep(self.P[f"feat_extractor.{i}.weight"], self.wp[i])
                if i in self.wpt:
                    be.conv_w_prep_t(self.P[f"feat_extractor.{i}.weight"], self.wpt[i])
            be.permute_cf(self.P["vgg2enc.weight"], self.vgg2enc_p, self.cfg.vgg_ch, self.cfg.f4, False)
        self.weights_dirty = False

    def load_flat(self, flat):
        """params <- flat (an arena of the same layout, e.g. the FOMAML meta weights): load_state_dict(_original) of
        run_task (fo_meta_interface.py:226) as one pass that also refreshes the compute-dtype shadow."""
        be = self.be
        n = self.layout.total
        if hasattr(be, "mt_copy_cast") and self.shadow is not None:
            if be.mt_copy_cast(self.params[:n], self.shadow[:n], flat[:n]):
                self.mark_shadow_fresh()
                return
        be.copy_(self.params, flat)
        self.weights_dirty = True

    # ------------------------------------------------------------------ helpers
    def site(self, name):
        if name not in self._sites:
            self._sites[name] = len(self._sites) + 1
        return self._sites[name]

    def _buf(self, ws, name, shape, dtype=None):
        t = ws.get(name)
        if t is None:
            t = torch.empty(shape, dtype=dtype or self.act_dtype, device=self.device)
            ws[name] = t
        return t

    def workspace(self, B, T, L1):
        key = (B, T, L1)
        if key not in self._ws:
            if len(self._ws) > 8:
                self._ws.clear()
            self._ws[key] = {}
        return self._ws[key]

    # ------------------------------------------------------------------ batch preparation (host)
    def prepare_batch(self, xs_pad, ilens, ys, olens=None):
        """Host-side bookkeeping of MyTransformer.forward/preprocess (:117,124-141): enc_lens,
        ys_in (sos + y, eos padded), ys_out (y + eos, IGNORE padded); olens += 1 in place."""
        cfg = self.cfg
        B = xs_pad.shape[0]
        assert B == ilens.shape[0] == len(ys), "Batch size mismatch"
        enc_lens = torch.floor(ilens.to(dtype=torch.float32) / 4).to(dtype=torch.int64)
        L1 = max(int(y.numel()) for y in ys) + 1
        # enc_lens | ys_in | ys_out live back to back in ONE int64 host buffer (pinned on a GPU box, taken from a
        # small ring): a single asynchronous H2D copy per batch.  Pageable sources would make every copy block
        # the host until the stream drains, which serialises the lanes / the run-ahead of the launch thread.
        ctc = cfg.ctc_weight > 0.0
        meta = self._host_meta(B * (1 + 2 * L1) + (2 * B if ctc else 0))
        meta[:B] = enc_lens
        ys_in = meta[B:B + B * L1].view(B, L1)
        ys_out = meta[B + B * L1:B + 2 * B * L1].view(B, L1)
        ys_in.fill_(cfg.eos_id)
        ys_out.fill_(IGNORE_ID)
        lens = [int(y.numel()) for y in ys]
        if min(lens) == L1 - 1:                       # equal-length batch (the bucketed train loader): three copies
            Y = torch.stack(list(ys))
            ys_in[:, 0] = cfg.sos_id
            ys_in[:, 1:] = Y
            ys_out[:, :L1 - 1] = Y
            ys_out[:, L1 - 1] = cfg.eos_id
        else:
            for b, y in enumerate(ys):
                n = lens[b]
                ys_in[b, 0] = cfg.sos_id
                ys_in[b, 1:n + 1] = y
                ys_out[b, :n] = y
                ys_out[b, n] = cfg.eos_id
        if olens is not None:
            olens += 1
        n_total = sum(lens) + B                       # non-pad targets = every y plus its eos
        if ctc:       # CTC reads its targets in place from ys_out (row b holds y_b first): per-utterance length / offset
            meta[B + 2 * B * L1:2 * B + 2 * B * L1] = torch.tensor(lens, dtype=torch.int64)
            meta[2 * B + 2 * B * L1:] = torch.arange(B, dtype=torch.int64) * L1
        hb = {"x": xs_pad, "meta": meta, "n_total": n_total, "B": B, "T": xs_pad.shape[1], "L1": L1}
        hb.update(self._meta_views(meta, B, L1))
        return hb

    def _host_meta(self, numel):
        """int64 host staging buffer from a ring of 64 (pinned when CUDA is present); an entry is reused only
        after the copy that last read it has completed (event per entry)."""
        ring = self.__dict__.get("_meta_ring")
        if ring is None:
            # ONE pinned allocation for the whole ring (cudaHostAlloc costs milliseconds: never on the step path)
            block = torch.empty(64 * 4096, dtype=torch.int64, pin_memory=self.device.type == "cuda")
            ring = self._meta_ring = {"i": 0, "bufs": [block[k * 4096:(k + 1) * 4096] for k in range(64)], "evs": [None] * 64}
        i = ring["i"] = (ring["i"] + 1) % 64
        if ring["evs"][i] is not None:
            ring["evs"][i].synchronize()
        t = ring["bufs"][i]
        if t.numel() < numel:                         # unusually large batch: private buffer for this entry
            t = torch.empty(numel, dtype=torch.int64, pin_memory=self.device.type == "cuda")
            ring["bufs"][i] = t
        return t[:numel]

    def _meta_copied(self, hb):
        """Record that the H2D copy of hb['meta'] has been enqueued on the current stream."""
        if self.device.type == "cuda" and "_meta_ring" in self.__dict__:
            ring = self._meta_ring
            for i, t in enumerate(ring["bufs"]):
                if t is not None and t.data_ptr() == hb["meta"].data_ptr():
                    ev = ring["evs"][i] or torch.cuda.Event()
                    ev.record(torch.cuda.current_stream(self.device))
                    ring["evs"][i] = ev
                    break

    def to_device(self, hb):
        """H2D of one prepared batch: x and the packed int64 block (two asynchronous copies)."""
        x = hb["x"]
        pinned_tmp = None
        if self.device.type == "cuda" and x.device.type == "cpu" and not x.is_pinned():
            x = pinned_tmp = x.pin_memory()        # (a loader with device= hands over features that are already resident)
        B, L1 = hb["B"], hb["L1"]
        meta = hb["meta"].to(self.device, non_blocking=True)
        self._meta_copied(hb)
        dev = {"x": x.to(self.device, non_blocking=True), "meta": meta}
        if pinned_tmp is not None:
            # the temporary pinned copy must outlive the asynchronous H2D copy that reads it: keep a reference until an
            # event recorded behind the copy has completed (belt and braces on top of torch's caching host allocator)
            held = self.__dict__.setdefault("_held_host", [])
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
            held.append((ev, pinned_tmp))
            while len(held) > 64 or (held and held[0][0].query()):
                if not held[0][0].query():
                    held[0][0].synchronize()
                held.pop(0)
        dev.update(self._meta_views(meta, B, L1))
        dev.update({k: hb[k] for k in ("n_total", "B", "T", "L1")})
        return dev

    def _meta_views(self, meta, B, L1):
        v = {"enc_lens": meta[:B], "ys_in": meta[B:B + B * L1].view(B, L1),
             "ys_out": meta[B + B * L1:B + 2 * B * L1].view(B, L1)}
        if self.cfg.ctc_weight > 0.0:
            v["ctc_tl"] = meta[B + 2 * B * L1:2 * B + 2 * B * L1]
            v["ctc_offs"] = meta[2 * B + 2 * B * L1:3 * B + 2 * B * L1]
        return v

    # ------------------------------------------------------------------ side stream (fork / join)
    def _fork(self, fn):
        """Run fn() on the side stream, ordered after everything queued so far on the current stream."""
        if not self.multi_stream:
            return fn()
        if self._side is None:
            self._side = torch.cuda.Stream(self.device)
        self._side.wait_stream(torch.cuda.current_stream(self.device))
        self.be.scratch_tag = (id(self), "side")
        with torch.cuda.stream(self._side):
            fn()
        self.be.scratch_tag = (id(self), "main")
        self._side_busy = True

    def _join(self):
        if self._side_busy:
            torch.cuda.current_stream(self.device).wait_stream(self._side)
            self._side_busy = False

    # ------------------------------------------------------------------ forward
    def forward(self, db, want_grad=True, mem=None):
        """Runs the network on a device batch; fills ws['logits'] [B*L1, C] fp32 and the loss
        statistics; with want_grad also d(mean loss)/d logits.  `mem` (greedy decode): an encoder memory
        [B*T', d] computed by an earlier call on the same x -- the conv front end and the encoder are skipped."""
        if db.get("ready") is not None:       # a batch staged on a copy stream (interfaces.stage_tasks)
            torch.cuda.current_stream(self.device).wait_event(db["ready"])
        if mem is None:
            self.forward_conv(db)
        return self.forward_rest(db, want_grad, mem)

    def _dims(self, db):
        cfg = self.cfg
        B, T, L1 = db["B"], db["T"], db["L1"]
        T2, F2 = T // 2, cfg.idim // 2
        T4, F4 = T2 // 2, F2 // 2
        ws = self.workspace(B, T, L1)
        ws["dims"] = (B, T, L1, T2, F2, T4, F4, B * T4, B * L1)
        return ws

    def forward_conv(self, db):
        """Segment 1 of a batch: derived weight copies + the VGG front end (NHWC) up to the second pool.  These are
        the long, GPU-filling launches; the lock-step scheduler of the meta-step queues them back to back for all
        task lanes (interfaces.FOMetaMixin._meta_lockstep)."""
        cfg, be = self.cfg, self.be
        self._bind_seed()
        self.prep_weights()
        ws = self._dims(db)
        B, T, L1, T2, F2, T4, F4, Me, Md = ws["dims"]
        F0 = cfg.idim
        P = self.P
        buf = lambda n, shape, dt=None: self._buf(ws, n, shape, dt)
        a1 = buf("a1", (B, T, F0, 64))
        be.conv1_fwd(db["x"], P["feat_extractor.0.weight"], P["feat_extractor.0.bias"], a1)
        a2 = buf("a2", (B, T, F0, 64))
        be.conv3x3_fwd(a1, self.wp[2], P["feat_extractor.2.bias"], a2)
        p1 = buf("p1", (B, T2, F2, 64))
        codes = getattr(be, "pool_codes", False)     # arg-max bytes: the pool backward does not re-read a2 / a4
        if codes:
            be.maxpool_fwd(a2, p1, code=buf("p1.code", (B, T2, F2, 64), torch.uint8))
        else:
            be.maxpool_fwd(a2, p1)
        a3 = buf("a3", (B, T2, F2, 128))
        be.conv3x3_fwd(p1, self.wp[5], P["feat_extractor.5.bias"], a3)
        a4 = buf("a4", (B, T2, F2, 128))
        be.conv3x3_fwd(a3, self.wp[7], P["feat_extractor.7.bias"], a4)
        p2 = buf("p2", (B, T4, F4, 128))
        if codes:
            be.maxpool_fwd(a4, p2, code=buf("p2.code", (B, T4, F4, 128), torch.uint8))
        else:
            be.maxpool_fwd(a4, p2)
        return ws

    def forward_rest(self, db, want_grad=True, mem=None):
        """Segment 2a: vgg2enc, encoder, decoder, output projection, loss (everything after the conv front end)."""
        cfg, be = self.cfg, self.be
        self._bind_seed()
        self.prep_weights()
        B, T, L1 = db["B"], db["T"], db["L1"]
        F0 = cfg.idim
        T2, F2 = T // 2, F0 // 2
        T4, F4 = T2 // 2, F2 // 2
        d, ff, H, C = cfg.d_model, cfg.d_inner, cfg.nheads, cfg.odim
        Me, Md = B * T4, B * L1
        ws = self._dims(db)
        buf = lambda n, shape, dt=None: self._buf(ws, n, shape, dt)
        f32 = torch.float32
        seed = self.step_seed
        pd = cfg.dropout if self.training else 0.0
        ppd = cfg.pos_dropout if self.training else 0.0
        P, W = self.P, self.W

        # ---- decoder prefix: the target embedding and layer 0's causal self-attention block do not depend on the
        # encoder at all -> side stream, concurrent with the conv front end / encoder
        def dec_self(l, x):
            pre = f"decoder.layers.{l}"
            ws[f"d{l}.in"] = x
            qkv = buf(f"d{l}.qkv", (Md, 3 * d))
            be.linear_fwd(x, W[pre + ".self_attn.in_proj_weight"], P[pre + ".self_attn.in_proj_bias"], qkv)
            ctx1 = buf(f"d{l}.ctx1", (Md, d))
            lse1 = buf(f"d{l}.lse1", (B * H * L1,), f32)
            be.attn_fwd(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], ctx1, lse1, B, H, L1, L1, None, True,
                        pd, seed, self.site(pre + ".sa"))
            s1 = buf(f"d{l}.s1", (Md, d))
            be.linear_fwd(ctx1, W[pre + ".self_attn.out_proj.weight"], P[pre + ".self_attn.out_proj.bias"], s1)
            h1 = buf(f"d{l}.h1", (Md, d))
            be.add_layernorm_fwd(s1, x, P[pre + ".norm1.weight"], P[pre + ".norm1.bias"], h1,
                                 buf(f"d{l}.m1", (Md,), f32), buf(f"d{l}.r1", (Md,), f32), pd, seed, self.site(pre + ".d1"))
            Wc, bc = W[pre + ".multihead_attn.in_proj_weight"], P[pre + ".multihead_attn.in_proj_bias"]
            q2 = buf(f"d{l}.q2", (Md, d))
            be.linear_fwd(h1, Wc[:d], bc[:d], q2)
            return h1, q2

        def kv_proj(l):
            pre = f"decoder.layers.{l}"
            Wc, bc = W[pre + ".multihead_attn.in_proj_weight"], P[pre + ".multihead_attn.in_proj_bias"]
            be.linear_fwd(ws["mem"], Wc[d:], bc[d:], buf(f"d{l}.kv2", (Me, 2 * d)))

        prefix = {}
        if cfg.dec_layers > 0:
            def dec_prefix():
                x0 = buf("d.x0", (Md, d))
                be.embed_pe_fwd(db["ys_in"].view(-1), P["pre_embed.weight"], self.pe2d, x0, L1, ppd, seed, self.site("dec.pe"))
                prefix["h1"], prefix["q2"] = dec_self(0, x0)
            self._fork(dec_prefix)

        mem_given = mem is not None
        if mem is not None:
            assert not want_grad, "a cached encoder memory is an inference-only shortcut"
            buf("mem", (Me, d)).copy_(mem)
        else:
            # ---- vgg2enc on the pooled conv output of forward_conv
            p2 = ws["p2"]
            h = buf("h0", (Me, d))
            be.linear_fwd(p2.view(Me, F4 * 128), self.vgg2enc_p, P["vgg2enc.bias"], h)
            be.add_pe_dropout(h, self.pe2d, T4, ppd, seed, self.site("enc.pe"))

            # ---- encoder (post-norm)
            for l in range(cfg.enc_layers):
                pre = f"encoder.layers.{l}"
                qkv = buf(f"e{l}.qkv", (Me, 3 * d))
                be.linear_fwd(h, W[pre + ".self_attn.in_proj_weight"], P[pre + ".self_attn.in_proj_bias"], qkv)
                ctx = buf(f"e{l}.ctx", (Me, d))
                lse = buf(f"e{l}.lse", (B * H * T4,), f32)
                be.attn_fwd(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], ctx, lse, B, H, T4, T4, db["enc_lens"], False,
                            pd, seed, self.site(pre + ".sa"))
                s1 = buf(f"e{l}.s1", (Me, d))
                be.linear_fwd(ctx, W[pre + ".self_attn.out_proj.weight"], P[pre + ".self_attn.out_proj.bias"], s1)
                h1 = buf(f"e{l}.h1", (Me, d))
                be.add_layernorm_fwd(s1, h, P[pre + ".norm1.weight"], P[pre + ".norm1.bias"], h1,
                                     buf(f"e{l}.m1", (Me,), f32), buf(f"e{l}.r1", (Me,), f32), pd, seed, self.site(pre + ".d1"))
                f1 = buf(f"e{l}.f1", (Me, ff))
                be.linear_fwd(h1, W[pre + ".linear1.weight"], P[pre + ".linear1.bias"], f1, relu=True,
                              dropout=(pd, seed, self.site(pre + ".df")))
                s2 = buf(f"e{l}.s2", (Me, d))
                be.linear_fwd(f1, W[pre + ".linear2.weight"], P[pre + ".linear2.bias"], s2)
                h2 = buf(f"e{l}.h2", (Me, d))
                be.add_layernorm_fwd(s2, h1, P[pre + ".norm2.weight"], P[pre + ".norm2.bias"], h2,
                                     buf(f"e{l}.m2", (Me,), f32), buf(f"e{l}.r2", (Me,), f32), pd, seed, self.site(pre + ".d2"))
                h = h2
            ws["enc_last"] = h
            mem = buf("mem", (Me, d))
            be.add_layernorm_fwd(h, None, P["encoder.norm.weight"], P["encoder.norm.bias"], mem,
                                 buf("enc.m", (Me,), f32), buf("enc.r", (Me,), f32), 0.0, seed, 0)

        ctc_here = cfg.ctc_weight > 0.0 and not mem_given
        # ---- decoder.  The memory K/V projections of layers 1.. only need `mem`: side stream, behind the prefix
        if cfg.dec_layers > 0:
            kv_proj(0)
            if cfg.dec_layers > 1:
                self._fork(lambda: [kv_proj(l) for l in range(1, cfg.dec_layers)])
        x = None
        for l in range(cfg.dec_layers):
            pre = f"decoder.layers.{l}"
            if l == 0:
                if cfg.dec_layers == 1:
                    self._join()
                h1, q2 = prefix["h1"], prefix["q2"]       # computed on the side stream (joined below / above)
            else:
                h1, q2 = dec_self(l, x)
            if l == 0 and cfg.dec_layers > 1:
                self._join()                              # prefix + K/V projections done (they were queued long ago)
            if l == 0 and ctc_here:
                self._fork(lambda: self._ctc_head(db, ws, want_grad))     # overlaps the rest of the decoder
                ctc_here = False
            kv2 = ws[f"d{l}.kv2"]
            ctx2 = buf(f"d{l}.ctx2", (Md, d))
            lse2 = buf(f"d{l}.lse2", (B * H * L1,), f32)
            be.attn_fwd(q2, kv2[:, :d], kv2[:, d:], ctx2, lse2, B, H, L1, T4, db["enc_lens"], False,
                        pd, seed, self.site(pre + ".ca"))
            s2 = buf(f"d{l}.s2", (Md, d))
            be.linear_fwd(ctx2, W[pre + ".multihead_attn.out_proj.weight"], P[pre + ".multihead_attn.out_proj.bias"], s2)
            h2 = buf(f"d{l}.h2", (Md, d))
            be.add_layernorm_fwd(s2, h1, P[pre + ".norm2.weight"], P[pre + ".norm2.bias"], h2,
                                 buf(f"d{l}.m2", (Md,), f32), buf(f"d{l}.r2", (Md,), f32), pd, seed, self.site(pre + ".d2"))
            f1 = buf(f"d{l}.f1", (Md, ff))
            be.linear_fwd(h2, W[pre + ".linear1.weight"], P[pre + ".linear1.bias"], f1, relu=True,
                          dropout=(pd, seed, self.site(pre + ".df")))
            s3 = buf(f"d{l}.s3", (Md, d))
            be.linear_fwd(f1, W[pre + ".linear2.weight"], P[pre + ".linear2.bias"], s3)
            h3 = buf(f"d{l}.h3", (Md, d))
            be.add_layernorm_fwd(s3, h2, P[pre + ".norm3.weight"], P[pre + ".norm3.bias"], h3,
                                 buf(f"d{l}.m3", (Md,), f32), buf(f"d{l}.r3", (Md,), f32), pd, seed, self.site(pre + ".d3"))
            x = h3
        if cfg.dec_layers == 0:
            x = buf("d.x0", (Md, d))
            be.embed_pe_fwd(db["ys_in"].view(-1), P["pre_embed.weight"], self.pe2d, x, L1, ppd, seed, self.site("dec.pe"))
        if ctc_here:
            self._ctc_head(db, ws, want_grad)
        ws["dec_last"] = x
        dout = buf("d.out", (Md, d))
        be.add_layernorm_fwd(x, None, P["decoder.norm.weight"], P["decoder.norm.bias"], dout,
                             buf("dec.m", (Md,), f32), buf("dec.r", (Md,), f32), 0.0, seed, 0)
        logits = buf("logits", (Md, C), f32)
        be.linear_fwd(dout, W["char_trans.weight"], P["char_trans.bias"], logits)

        # ---- label-smoothed CE, accuracy, d logits
        be.zero_(self.stats)
        argmax = buf("argmax", (Md,), torch.int64)
        # d logits in the compute dtype, rows padded to a multiple of 8 elements so that (in bf16 mode) the
        # output-projection dgrad / wgrad run on the tcgen05 path (TMA needs 16-byte row strides)
        dlogits = None
        if want_grad:
            dlogits = buf("dlogits", (Md, (C + 7) // 8 * 8))[:, :C]
            ws["dlogits_v"] = dlogits
        w = cfg.ctc_weight
        be.ls_ce(logits, db["ys_out"].view(-1), self.eps_ls, (1.0 - w) / max(db["n_total"], 1), self.stats, argmax, dlogits,
                 db.get("inv_n_dev"))
        if w > 0.0 and not mem_given:
            self._join()
            be.loss_mix(self.stats, ws["ctc.loss"], w)
        return ws

    def _ctc_head(self, db, ws, want_grad):
        """CTC branch of the joint objective: logits = ctc_lo(encoder memory) [B*T', C] fp32, fused log-softmax + alpha-beta
        forward-backward (kernel 1) reading the batch-first rows in place; the gradient leaves already scaled by w."""
        cfg, be = self.cfg, self.be
        B, T, L1, T2, F2, T4, F4, Me, Md = ws["dims"]
        C = cfg.odim
        f32 = torch.float32
        cl = self._buf(ws, "ctc.logits", (Me, C), f32)
        be.linear_fwd(ws["mem"], self.W["ctc_lo.weight"], self.P["ctc_lo.bias"], cl)
        g = self._buf(ws, "ctc.grad", (Me, C), f32) if want_grad else None
        be.ctc_joint(cl, B, T4, C, db["ys_out"].view(-1), db["ctc_offs"], db["enc_lens"], db["ctc_tl"], L1 - 1,
                     cfg.ctc_weight, self._buf(ws, "ctc.nll", (B,), f32), self._buf(ws, "ctc.loss", (1,), f32), g)

    # ------------------------------------------------------------------ backward
    def backward(self, db, ws):
        """Back-propagates ws['dlogits'] through the whole network; gradients are ACCUMULATED into
        self.grads (zeroed here first, as run_batch's asr_opt.zero_grad() does)."""
        self.backward_rest(db, ws, join=False)
        self.backward_conv(db, ws)

    def backward_rest(self, db, ws, join=True):
        """Segment 2b: output projection, decoder, encoder and vgg2enc backward, down to the gradient of the pooled conv
        output (ws['g.p2'])."""
        cfg, be = self.cfg, self.be
        self._bind_seed()
        B, T, L1, T2, F2, T4, F4, Me, Md = ws["dims"]
        F0 = cfg.idim
        d, ff, H, C = cfg.d_model, cfg.d_inner, cfg.nheads, cfg.odim
        buf = lambda n, shape, dt=None: self._buf(ws, n, shape, dt)
        f32 = torch.float32
        seed = self.step_seed
        pd = cfg.dropout if self.training else 0.0
        ppd = cfg.pos_dropout if self.training else 0.0
        P, W, G = self.P, self.W, self.G
        be.zero_(self.grads)

        fork = self._fork

        # Weight gradients never feed the dgrad chain.  Where the backend can group launches (tcgen05 path) they are
        # collected and issued as ONE grouped launch per stretch (decoder, encoder) on the side stream: ~26 problems of
        # 16-64 CTAs each otherwise.  Every operand of a deferred problem is a forward activation or a gradient buffer
        # private to its (layer, sub-layer), so nothing overwrites it before the flush.
        grouped = self.group_wgrads and hasattr(be, "gemm_group") and getattr(be, "gemm_path", "") == "umma" and \
            self.act_dtype != torch.float32
        pending = []

        def wgrad(x, dy, dw, db):
            if grouped:
                pending.append(lambda: be.linear_wgrad(x, dy, dw, db))
            else:
                fork(lambda: be.linear_wgrad(x, dy, dw, db))

        def flush_wgrads():
            if pending:
                calls = list(pending)
                pending.clear()
                fork(lambda: be.gemm_group(calls))

        def ffn_bwd(pre, g_s, f1, h_in, g_res, Mrows, tag):
            """g_s: grad wrt linear2 output; accumulates the FFN input gradient into g_res."""
            wgrad(f1, g_s, G[pre + ".linear2.weight"], G[pre + ".linear2.bias"])
            g_f1 = buf(tag + ".g_f1", (Mrows, ff))
            # f1 = dropout(relu(.)): its backward (mask f1 > 0, scale 1/(1-p)) is fused into the dgrad epilogue
            be.linear_dgrad(g_s, W[pre + ".linear2.weight"], g_f1, relu_drop_mask=f1, p=pd)
            wgrad(h_in, g_f1, G[pre + ".linear1.weight"], G[pre + ".linear1.bias"])
            be.linear_dgrad(g_f1, W[pre + ".linear1.weight"], g_res, accumulate=True)

        def self_attn_bwd(pre, g_o, qkv, ctx, lse, x_in, g_res, Mrows, L, klens, causal, tag, site):
            wgrad(ctx, g_o, G[pre + ".self_attn.out_proj.weight"], G[pre + ".self_attn.out_proj.bias"])
            g_ctx = buf(tag[:2] + ".g_ctx", (Mrows, d))
            dsum = buf(tag[:2] + ".dsum", (B * H * L,), f32)
            # D = rowsum(dO . O) per head comes out of the out-projection dgrad's epilogue (no separate kernel)
            ready = be.linear_dgrad(g_o, W[pre + ".self_attn.out_proj.weight"], g_ctx, rowdot=(ctx, dsum, L, H))
            g_qkv = buf(tag + ".g_qkv", (Mrows, 3 * d))
            be.attn_bwd(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], ctx, g_ctx, lse, dsum,
                        g_qkv[:, :d], g_qkv[:, d:2 * d], g_qkv[:, 2 * d:], B, H, L, L, klens, causal, pd, seed, site,
                        dsum_ready=bool(ready))
            wgrad(x_in, g_qkv, G[pre + ".self_attn.in_proj_weight"], G[pre + ".self_attn.in_proj_bias"])
            be.linear_dgrad(g_qkv, W[pre + ".self_attn.in_proj_weight"], g_res, accumulate=True)

        # Buffers read by a forked weight-gradient GEMM are private to their (layer, sub-layer): the dgrad
        # chain on the main stream must not overwrite them while the side stream still reads them.
        # ---- output projection + final decoder norm
        dlogits = ws["dlogits_v"]
        wgrad(ws["d.out"], dlogits, G["char_trans.weight"], G["char_trans.bias"])
        g_out = buf("g.out", (Md, d))
        be.linear_dgrad(dlogits, W["char_trans.weight"], g_out)
        gd = [buf("g.dA", (Md, d)), buf("g.dB", (Md, d))]
        cur = 0
        be.add_layernorm_bwd(g_out, ws["dec_last"], ws["dec.m"], ws["dec.r"], P["decoder.norm.weight"], gd[cur], False,
                             None, G["decoder.norm.weight"], G["decoder.norm.bias"])
        g_mem = buf("g.mem", (Me, d))
        first_mem = True
        for l in reversed(range(cfg.dec_layers)):
            pre = f"decoder.layers.{l}"
            tag = f"gd{l}"
            # norm3 / FFN
            nxt = 1 - cur
            g_br = buf(tag + ".br3", (Md, d))                  # gradient of the sub-layer (branch) output
            be.add_layernorm_bwd(gd[cur], ws[f"d{l}.s3"], ws[f"d{l}.m3"], ws[f"d{l}.r3"], P[pre + ".norm3.weight"],
                                 gd[nxt], False, g_br, G[pre + ".norm3.weight"], G[pre + ".norm3.bias"],
                                 pd, seed, self.site(pre + ".d3"))
            ffn_bwd(pre, g_br, ws[f"d{l}.f1"], ws[f"d{l}.h2"], gd[nxt], Md, tag)
            cur = nxt
            # norm2 / cross attention
            nxt = 1 - cur
            g_br = buf(tag + ".br2", (Md, d))
            be.add_layernorm_bwd(gd[cur], ws[f"d{l}.s2"], ws[f"d{l}.m2"], ws[f"d{l}.r2"], P[pre + ".norm2.weight"],
                                 gd[nxt], False, g_br, G[pre + ".norm2.weight"], G[pre + ".norm2.bias"],
                                 pd, seed, self.site(pre + ".d2"))
            wgrad(ws[f"d{l}.ctx2"], g_br, G[pre + ".multihead_attn.out_proj.weight"],
                  G[pre + ".multihead_attn.out_proj.bias"])
            g_ctx = buf("gd.g_ctx", (Md, d))
            dsum = buf("gd.dsum", (B * H * L1,), f32)
            ready = be.linear_dgrad(g_br, W[pre + ".multihead_attn.out_proj.weight"], g_ctx,
                                    rowdot=(ws[f"d{l}.ctx2"], dsum, L1, H))
            g_q2 = buf(tag + ".g_q2", (Md, d))
            g_kv2 = buf(tag + ".g_kv2", (Me, 2 * d))
            kv2 = ws[f"d{l}.kv2"]
            be.attn_bwd(ws[f"d{l}.q2"], kv2[:, :d], kv2[:, d:], ws[f"d{l}.ctx2"], g_ctx, ws[f"d{l}.lse2"], dsum,
                        g_q2, g_kv2[:, :d], g_kv2[:, d:], B, H, L1, T4, db["enc_lens"], False,
                        pd, seed, self.site(pre + ".ca"), dsum_ready=bool(ready))
            Wc = W[pre + ".multihead_attn.in_proj_weight"]
            Gw, Gb = G[pre + ".multihead_attn.in_proj_weight"], G[pre + ".multihead_attn.in_proj_bias"]
            wgrad(ws[f"d{l}.h1"], g_q2, Gw[:d], Gb[:d])
            be.linear_dgrad(g_q2, Wc[:d], gd[nxt], accumulate=True)
            wgrad(ws["mem"], g_kv2, Gw[d:], Gb[d:])
            be.linear_dgrad(g_kv2, Wc[d:], g_mem, accumulate=not first_mem)
            first_mem = False
            cur = nxt
            # norm1 / causal self attention
            nxt = 1 - cur
            g_br = buf(tag + ".br1", (Md, d))
            be.add_layernorm_bwd(gd[cur], ws[f"d{l}.s1"], ws[f"d{l}.m1"], ws[f"d{l}.r1"], P[pre + ".norm1.weight"],
                                 gd[nxt], False, g_br, G[pre + ".norm1.weight"], G[pre + ".norm1.bias"],
                                 pd, seed, self.site(pre + ".d1"))
            self_attn_bwd(pre, g_br, ws[f"d{l}.qkv"], ws[f"d{l}.ctx1"], ws[f"d{l}.lse1"], ws[f"d{l}.in"], gd[nxt],
                          Md, L1, None, True, tag, self.site(pre + ".sa"))
            cur = nxt
        flush_wgrads()                 # the decoder's weight gradients: one grouped launch under the encoder's dgrad chain
        # the scatter-add into the tied embedding matrix shares its target with the char_trans wgrad: keep
        # both on the side stream (ordered); nothing on the main stream touches gd[] after this point
        g_emb = gd[cur]
        fork(lambda: be.embed_bwd(db["ys_in"].view(-1), g_emb, G["pre_embed.weight"], L1, ppd, seed, self.site("dec.pe")))

        # ---- encoder
        ge = [buf("g.eA", (Me, d)), buf("g.eB", (Me, d))]
        cur = 0
        if cfg.dec_layers == 0:
            be.zero_(g_mem)
        if cfg.ctc_weight > 0.0:
            gc = ws["ctc.grad"]
            if self.act_dtype != torch.float32:       # bf16 rows padded to a 16-byte pitch for the tcgen05 GEMMs
                C8 = (C + 7) // 8 * 8
                gcb = buf("ctc.grad_c", (Me, C8))
                be.cast_pad2d(gc, gcb, C)
                gc = gcb[:, :C]
            wgrad(ws["mem"], gc, G["ctc_lo.weight"], G["ctc_lo.bias"])
            be.linear_dgrad(gc, W["ctc_lo.weight"], g_mem, accumulate=True)
        be.add_layernorm_bwd(g_mem, ws["enc_last"], ws["enc.m"], ws["enc.r"], P["encoder.norm.weight"], ge[cur], False,
                             None, G["encoder.norm.weight"], G["encoder.norm.bias"])
        for l in reversed(range(cfg.enc_layers)):
            pre = f"encoder.layers.{l}"
            tag = f"ge{l}"
            nxt = 1 - cur
            g_ebr = buf(tag + ".br2", (Me, d))
            be.add_layernorm_bwd(ge[cur], ws[f"e{l}.s2"], ws[f"e{l}.m2"], ws[f"e{l}.r2"], P[pre + ".norm2.weight"],
                                 ge[nxt], False, g_ebr, G[pre + ".norm2.weight"], G[pre + ".norm2.bias"],
                                 pd, seed, self.site(pre + ".d2"))
            ffn_bwd(pre, g_ebr, ws[f"e{l}.f1"], ws[f"e{l}.h1"], ge[nxt], Me, tag)
            cur = nxt
            nxt = 1 - cur
            g_ebr = buf(tag + ".br1", (Me, d))
            be.add_layernorm_bwd(ge[cur], ws[f"e{l}.s1"], ws[f"e{l}.m1"], ws[f"e{l}.r1"], P[pre + ".norm1.weight"],
                                 ge[nxt], False, g_ebr, G[pre + ".norm1.weight"], G[pre + ".norm1.bias"],
                                 pd, seed, self.site(pre + ".d1"))
            x_in = ws["h0"] if l == 0 else ws[f"e{l - 1}.h2"]
            self_attn_bwd(pre, g_ebr, ws[f"e{l}.qkv"], ws[f"e{l}.ctx"], ws[f"e{l}.lse"], x_in, ge[nxt],
                          Me, T4, db["enc_lens"], False, tag, self.site(pre + ".sa"))
            cur = nxt
        flush_wgrads()                 # the encoder's
        g_h0 = ge[cur]
        be.dropout(g_h0, ppd, seed, self.site("enc.pe"))

        # ---- vgg2enc + VGG front end
        p2f = ws["p2"].view(Me, F4 * 128)

        def vgg2enc_wgrad():
            be.zero_(self.d_vgg2enc_p)
            be.linear_wgrad(p2f, g_h0, self.d_vgg2enc_p, G["vgg2enc.bias"])
            be.permute_cf(self.d_vgg2enc_p, G["vgg2enc.weight"], cfg.vgg_ch, cfg.f4, True)
        fork(vgg2enc_wgrad)
        g_p2 = buf("g.p2", (B, T4, F4, 128))
        be.linear_dgrad(g_h0, self.vgg2enc_p, g_p2.view(Me, F4 * 128))
        if join:                  # a segment that is captured / scheduled on its own must end with its side work joined
            self._join()

    def backward_conv(self, db, ws):
        """Segment 3: VGG front end backward (pool / dgrad chain on the stream, weight gradients on the side stream)."""
        cfg, be = self.cfg, self.be
        self._bind_seed()
        B, T, L1, T2, F2, T4, F4, Me, Md = ws["dims"]
        F0 = cfg.idim
        buf = lambda n, shape, dt=None: self._buf(ws, n, shape, dt)
        G = self.G
        fork = self._fork
        g_p2 = ws["g.p2"]
        g_a4 = buf("g.a4", (B, T2, F2, 128))
        pool_kw = lambda nm: {"code": ws[nm]} if nm in ws else {}
        be.maxpool_bwd(ws["a4"], g_p2, g_a4, True, **pool_kw("p2.code"))

        def conv_wgrad(i, x, dy):
            def run():
                be.zero_(self.dwp[i])
                be.conv3x3_wgrad(x, dy, self.dwp[i], G[f"feat_extractor.{i}.bias"])
                be.conv_w_unprep_add(self.dwp[i], G[f"feat_extractor.{i}.weight"])
            fork(run)

        conv_wgrad(7, ws["a3"], g_a4)
        g_a3 = buf("g.a3", (B, T2, F2, 128))
        be.conv3x3_dgrad(g_a4, self.wp[7], g_a3, ws["a3"], **({"wpt": self.wpt[7]} if 7 in self.wpt else {}))
        conv_wgrad(5, ws["p1"], g_a3)
        g_p1 = buf("g.p1", (B, T2, F2, 64))
        be.conv3x3_dgrad(g_a3, self.wp[5], g_p1, None, **({"wpt": self.wpt[5]} if 5 in self.wpt else {}))
        g_a2 = buf("g.a2", (B, T, F0, 64))
        be.maxpool_bwd(ws["a2"], g_p1, g_a2, True, **pool_kw("p1.code"))
        conv_wgrad(2, ws["a1"], g_a2)
        g_a1 = buf("g.a1", (B, T, F0, 64))
        be.conv3x3_dgrad(g_a2, self.wp[2], g_a1, ws["a1"], **({"wpt": self.wpt[2]} if 2 in self.wpt else {}))
        be.conv1_wgrad(db["x"], g_a1, G["feat_extractor.0.weight"], G["feat_extractor.0.bias"])
        self._join()

    # ------------------------------------------------------------------ greedy decoding with a key/value cache
    @torch.no_grad()
    def greedy_decode(self, xs_pad, ilens, n_steps):
        """MyTransformer.recog (mono_transformer_torch.py:143-176) without its O(L^2) decoder re-runs.  The reference
        feeds the growing prefix [sos, y_0 .. y_{j-1}] through the whole decoder at step j and re-decides EVERY position;
        the decoder is causal, so positions < j reproduce their earlier decision and only position j is new.  Here step
        0 is the ordinary forward on the prefix [sos] (conv front end, encoder, memory K/V projections of every layer);
        every later step runs ONE decoder row per utterance against the cached self-attention keys / values (fixed
        capacity, read in place by the attention kernel) and the memory K/V of step 0.  Returns ids [n_steps, B]."""
        cfg, be = self.cfg, self.be
        B = xs_pad.shape[0]
        d, ff, H, C = cfg.d_model, cfg.d_inner, cfg.nheads, cfg.odim
        dev, adt, f32 = self.device, self.act_dtype, torch.float32
        was_training, self.training = self.training, False
        self.weights_dirty = True
        empty = [torch.zeros(0, dtype=torch.int64) for _ in range(B)]
        db = self.to_device(self.prepare_batch(xs_pad, ilens, empty, None))
        ws0 = self.forward(db, want_grad=False)
        ids = [ws0["argmax"].view(B).clone()]
        if n_steps > 1:
            T4 = ws0["dims"][5]
            cap = n_steps
            P, W = self.P, self.W
            kc = [torch.zeros(B * cap, d, dtype=adt, device=dev) for _ in range(cfg.dec_layers)]
            vc = [torch.zeros(B * cap, d, dtype=adt, device=dev) for _ in range(cfg.dec_layers)]
            for l in range(cfg.dec_layers):
                qkv0 = ws0[f"d{l}.qkv"]
                kc[l].view(B, cap, d)[:, 0].copy_(qkv0[:, d:2 * d])
                vc[l].view(B, cap, d)[:, 0].copy_(qkv0[:, 2 * d:])
            mk = lambda *shape, dt=adt: torch.empty(*shape, dtype=dt, device=dev)
            x0, qkv, ctx, s, q2 = mk(B, d), mk(B, 3 * d), mk(B, d), mk(B, d), mk(B, d)
            h1, h2, h3a, h3b, f1, dout = mk(B, d), mk(B, d), mk(B, d), mk(B, d), mk(B, ff), mk(B, d)
            lse, mean, rstd = mk(B * H, dt=f32), mk(B, dt=f32), mk(B, dt=f32)
            logits, argmax = mk(B, C, dt=f32), mk(B, dt=torch.int64)
            gold = torch.full((B,), IGNORE_ID, dtype=torch.int64, device=dev)
            self._bind_seed()
            for j in range(1, n_steps):
                be.embed_pe_fwd(ids[-1], P["pre_embed.weight"], self.pe2d[j:j + 1], x0, 1, 0.0, 0, 0)
                x = x0
                for l in range(cfg.dec_layers):
                    pre = f"decoder.layers.{l}"
                    be.linear_fwd(x, W[pre + ".self_attn.in_proj_weight"], P[pre + ".self_attn.in_proj_bias"], qkv)
                    kc[l].view(B, cap, d)[:, j].copy_(qkv[:, d:2 * d])
                    vc[l].view(B, cap, d)[:, j].copy_(qkv[:, 2 * d:])
                    be.attn_fwd(qkv[:, :d], kc[l], vc[l], ctx, lse, B, H, 1, j + 1, None, False, kv_rows=cap)
                    be.linear_fwd(ctx, W[pre + ".self_attn.out_proj.weight"], P[pre + ".self_attn.out_proj.bias"], s)
                    be.add_layernorm_fwd(s, x, P[pre + ".norm1.weight"], P[pre + ".norm1.bias"], h1, mean, rstd, 0.0, 0, 0)
                    Wc, bc = W[pre + ".multihead_attn.in_proj_weight"], P[pre + ".multihead_attn.in_proj_bias"]
                    be.linear_fwd(h1, Wc[:d], bc[:d], q2)
                    kv2 = ws0[f"d{l}.kv2"]
                    be.attn_fwd(q2, kv2[:, :d], kv2[:, d:], ctx, lse, B, H, 1, T4, db["enc_lens"], False)
                    be.linear_fwd(ctx, W[pre + ".multihead_attn.out_proj.weight"], P[pre + ".multihead_attn.out_proj.bias"], s)
                    be.add_layernorm_fwd(s, h1, P[pre + ".norm2.weight"], P[pre + ".norm2.bias"], h2, mean, rstd, 0.0, 0, 0)
                    be.linear_fwd(h2, W[pre + ".linear1.weight"], P[pre + ".linear1.bias"], f1, relu=True)
                    be.linear_fwd(f1, W[pre + ".linear2.weight"], P[pre + ".linear2.bias"], s)
                    h3 = h3b if x is h3a else h3a               # the layer's input is still its first residual
                    be.add_layernorm_fwd(s, h2, P[pre + ".norm3.weight"], P[pre + ".norm3.bias"], h3, mean, rstd, 0.0, 0, 0)
                    x = h3
                be.add_layernorm_fwd(x, None, P["decoder.norm.weight"], P["decoder.norm.bias"], dout, mean, rstd, 0.0, 0, 0)
                be.linear_fwd(dout, W["char_trans.weight"], P["char_trans.bias"], logits)
                be.zero_(self.stats)
                be.ls_ce(logits, gold, self.eps_ls, 1.0, self.stats, argmax, None, None)
                ids.append(argmax.clone())
        self.training = was_training
        return torch.stack(ids, 0)

    # ------------------------------------------------------------------ public step
    N_SEGMENTS = 3

    def forward_backward(self, db):
        """One run_batch(train=True) worth of device work.  Returns the workspace; the loss
        statistics stay on the device in self.stats = [sum of row losses, n_correct, n_non_pad, ...].
        The work is issued as THREE segments -- conv front end forward | everything between | conv front end backward
        -- here back to back on the current stream; the lock-step meta-step scheduler issues the segments of several
        task lanes on different streams (fb_begin / fb_segment).  With use_graphs each segment of a (B, T, L1) shape is
        captured once into a CUDA graph and replayed: inputs are copied into static buffers, the dropout seed offset
        and 1/n live in device memory, so a batch is three launches instead of ~240."""
        h = self.fb_begin(db)
        for i in range(self.N_SEGMENTS):
            self.fb_segment(h, i)
        return h["ws"] if h["ws"] is not None else self.workspace(db["B"], db["T"], db["L1"])

    def fb_begin(self, db):
        """Stages the inputs of a batch on the CURRENT stream and returns the handle fb_segment takes."""
        if db.get("ready") is not None:       # staged ahead of time on a copy stream (interfaces.stage_tasks)
            torch.cuda.current_stream(self.device).wait_event(db["ready"])
        if self.use_graphs and self.device.type == "cuda":
            ent = self._graph_entry(db)
            graphs, sdb, ws = ent
            self._load_static(sdb, db)
            # the derived weight copies are refreshed OUTSIDE the captured segments (one launch): what it has to do
            # depends on who moved the weights (an optimizer kernel that wrote the shadow already, or anything else)
            self.weights_dirty = True
            self.prep_weights()
            return {"graphs": graphs, "db": sdb, "ws": ws}
        return {"graphs": None, "db": db, "ws": None}

    def fb_segment(self, h, i):
        """Issues segment i (0: conv forward, 1: the rest of forward + backward down to the conv output, 2: conv
        backward) of the batch on the current stream."""
        if h["graphs"] is not None:
            h["graphs"][i].replay()
        else:
            self._segment_eager(h["db"], i, join=True)

    def _bind_seed(self):
        self.be.scratch_tag = (id(self), "main")   # backend scratch buffers are private to this engine
        if hasattr(self.be, "set_seed_ptr"):
            self.be.set_seed_ptr(self.seed_t)  # kernels launched from here on read this engine's offset

    def _segment_eager(self, db, i, join, prep=True):
        if i == 0:
            if prep:
                self.weights_dirty = True     # training: the master weights may have moved since the last batch
            self._bind_seed()
            if hasattr(self.be, "seed_bump"):
                self.be.seed_bump(1)          # fresh dropout masks: device-resident seed offset += 1
            else:
                self.step_seed += 1
            self.forward_conv(db)
        elif i == 1:
            ws = self.forward_rest(db, want_grad=True)
            self.backward_rest(db, ws, join=join)
        else:
            self.backward_conv(db, self.workspace(db["B"], db["T"], db["L1"]))

    def _forward_backward_eager(self, db):
        for i in range(self.N_SEGMENTS):
            self._segment_eager(db, i, join=False)
        return self.workspace(db["B"], db["T"], db["L1"])

    def _graph_entry(self, db):
        B, T, L1 = db["B"], db["T"], db["L1"]
        key = (B, T, L1, self.training, bool(self.group_wgrads))
        ent = self._graphs.get(key)
        if ent is not None:
            self._graphs.move_to_end(key)
            return ent
        # bounded LRU: with the reference's 1-frame buckets nearly every real batch has a new (B, T, L) shape; a
        # captured graph pins its workspaces (hundreds of MB at full size), so keep only the most recent shapes
        while len(self._graphs) >= self.max_graphs:
            old_key, _ = self._graphs.popitem(last=False)
            if not any(k[:3] == old_key[:3] for k in self._graphs):      # (another variant of the shape may still use it)
                self._ws.pop(old_key[:3], None)
        dev = self.device
        smeta = torch.empty(B * (1 + 2 * L1) + (2 * B if self.cfg.ctc_weight > 0.0 else 0), dtype=torch.int64, device=dev)
        sdb = {"x": torch.empty(B, T, self.cfg.idim, dtype=torch.float32, device=dev), "meta": smeta,
               "inv_n_dev": torch.empty(1, dtype=torch.float32, device=dev),
               "n_total": 1, "B": B, "T": T, "L1": L1}
        sdb.update(self._meta_views(smeta, B, L1))
        self._load_static(sdb, db)
        cur = torch.cuda.current_stream(dev)
        side = torch.cuda.Stream(dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side):         # allocate workspaces / set kernel attributes eagerly first
            self._forward_backward_eager(sdb)
        cur.wait_stream(side)
        torch.cuda.synchronize(dev)
        graphs = []
        for i in range(self.N_SEGMENTS):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._segment_eager(sdb, i, join=True, prep=False)      # fb_begin preps eagerly before the replays
            graphs.append(g)
        ent = (graphs, sdb, self.workspace(B, T, L1))
        self._graphs[key] = ent
        return ent

    def _load_static(self, sdb, db):
        sdb["x"].copy_(db["x"], non_blocking=True)
        if "meta" in db:
            sdb["meta"].copy_(db["meta"], non_blocking=True)
            if db["meta"].device.type == "cpu":
                self._meta_copied(db)
        else:
            for k in ("enc_lens", "ys_in", "ys_out", "ctc_tl", "ctc_offs"):
                if k in sdb:
                    sdb[k].copy_(db[k], non_blocking=True)
        sdb["inv_n_dev"].fill_((1.0 - self.cfg.ctc_weight) / max(db["n_total"], 1))

    def read_stats(self):
        """The single device->host read of a step: {'loss', 'acc'} like run_batch's info dict."""
        return self.stats_to_info(self.stats.tolist())

    @staticmethod
    def stats_to_info(s):
        """[sum of row losses, n_correct, n_non_pad, CTC term, w, ...] -> {'loss','acc'}; w = 0 unless the joint
        CTC / attention objective is on, in which case loss = (1-w) * LS-CE + w * CTC and both terms are reported."""
        n = max(s[2], 1.0)
        att = s[0] / n
        w = s[4] if len(s) > 4 else 0.0
        if w > 0.0:
            return {"loss": (1.0 - w) * att + w * s[3], "acc": s[1] / n, "att_loss": att, "ctc_loss": s[3]}
        return {"loss": att, "acc": s[1] / n}
