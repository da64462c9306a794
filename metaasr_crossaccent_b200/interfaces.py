"""Fused meta-learning interfaces: drop-ins for the reference's training loops

    FOMetaASRInterface  (src/fo_meta_interface.py:18-302)   FOMAML  (+ Reptile, which the
                         reference accepts on the CLI but raises on, :195-198 -- SURVEY 8a row R)
    MultiASRInterface   (src/multi_interface.py:17-192)     multi-task
    MonoASRInterface    (src/mono_interface.py:18-230)      fine-tune / mono-accent loop (filter_model,
                         freeze_module, per-epoch checkpoint) -- SURVEY 8f #2

They keep the reference's hook names and behaviour (`load_model`, `run_task`,
`_partial_meta_update`, `_final_meta_update`, `train`, `save_per_steps`, `_original`,
`_updates`, `_counter`, `meta_opt`, `inner_lr`) but run every parameter-space step as ONE fused
kernel over the flat arenas of engine.py and combine the ranks' meta-gradients with a single
all-reduce (dist.py).  Two ways to use them:

  * inside the reference tree: `fused(FOMetaMixin, src.fo_meta_interface.FOMetaASRInterface)`
    keeps the reference's __init__/logging/eval/checkpoint code and swaps only the hot hooks
    (INTEGRATION.md);
  * standalone (bench, tests, GPU box without the reference): FOMetaASRInterface /
    MultiASRInterface below, built on the small PretrainHost base.
"""
from __future__ import annotations

import math
import pickle
import random
from collections import OrderedDict
from functools import partial
from pathlib import Path

import torch

from . import dist as D
from .optim import FlatAdamState, FlatInnerSGD, FlatNoamAdam, noam_lr

GRAD_CLIP = 5                  # src/marcos.py:7
LOG_DIR = 'testing-logs'       # src/marcos.py:4
INIT_BEST_ER = 200.0           # src/marcos.py:5
SOS_SYMBOL, EOS_SYMBOL = '<s>', '</s>'


class RunningAvgDict(dict):
    """Stand-in for torchexp.stat.RunningAvgDict (pretrain_interface.py:12): decay 1.0 ->
    count-weighted mean, else exponential moving average.  Picklable (info_dict.latest)."""

    def __init__(self, decay_rate=0.99):
        super().__init__()
        self.decay_rate = decay_rate
        self._n = {}

    def add(self, info, n=1):
        for k, v in info.items():
            v = float(v)
            if k not in self:
                self[k], self._n[k] = v, n
            elif self.decay_rate >= 1.0:
                tot = self._n[k] + n
                self[k] = (self[k] * self._n[k] + v * n) / tot
                self._n[k] = tot
            else:
                self[k] = self.decay_rate * self[k] + (1 - self.decay_rate) * v


def _make_metric(solver_cfg, units):
    """Metric(spm_model, id2units, sos, eos) as in train_interface.py:37-41 / pretrain_interface.py:43-48 when
    `solver.spm_mapping` was readable (real vocabulary); synthetic runs score nothing (cer / wer = nan)."""
    if not Path(solver_cfg.get('spm_mapping', '')).is_file():
        return None
    from .metric import Metric
    return Metric(solver_cfg.get('spm_model'), units, 0, len(units) - 1)


class _NullDashboard:
    def __getattr__(self, name):
        return lambda *a, **k: None


class PretrainHost:
    """Minimal stand-alone equivalent of PretrainInterface (src/pretrain_interface.py:16-138):
    config fields, vocabulary (<s> + units + </s>), log-dir layout, text logs.  Data comes from
    `self.data_container` (anything with get_item(accent_idx=None, num=1), io/dataset.py:248)."""

    def __init__(self, config, paras, id2accent):
        self.config, self.paras = config, paras
        self.train_type = 'pretrain'
        s = config['solver']
        self.eval_ival, self.log_ival, self.save_ival = s['eval_ival'], s['log_ival'], s['save_ival']
        self.half_batch_ilen, self.dev_max_ilen = s.get('half_batch_ilen'), s.get('dev_max_ilen')
        self.sample_strategy = getattr(paras, 'sample_strategy', 'normal')
        self.best_cer = self.best_wer = INIT_BEST_ER
        units = [SOS_SYMBOL]
        mapping = Path(s.get('spm_mapping', ''))
        if mapping.is_file():
            with open(mapping) as fin:
                units += [line.rstrip().split(' ')[0] for line in fin.readlines()]
        else:                                     # synthetic runs: unigram150 has 365 units -> odim 367
            units += [f"<u{i}>" for i in range(1, int(s.get('n_units', 365)) + 1)]
        units.append(EOS_SYMBOL)
        self.id2units = units
        self.id2ch = units
        self.metric_observer = _make_metric(s, units)   # CER / WER scoring when the sentencepiece files are configured
        self.accents = [id2accent[a] for a in paras.pretrain_accents]
        self.num_pretrain = paras.num_pretrain
        self.tgt_accent = id2accent.get(getattr(paras, 'tgt_accent', None), 'none')
        self.max_step = paras.max_step if paras.max_step > 0 else s['total_steps']
        assert self.num_pretrain == len(self.accents)
        self.log_dir = None
        if getattr(paras, 'log_root', None) is not None and D.rank() == 0:
            self.log_dir = Path(paras.log_root, LOG_DIR, self.train_type, s['setting'], paras.algo,
                                paras.pretrain_suffix, self.tgt_accent, str(paras.runs))
            self.log_dir.mkdir(parents=True, exist_ok=True)
        self.train_info = RunningAvgDict(decay_rate=0.99)
        self.global_step = 1
        if getattr(paras, 'resume', False):
            # pretrain_interface.py:82-103.  (The reference asserts an optimizer.latest that none of its pretraining
            # loops writes; save_per_steps below writes one, so resuming works here.)
            root = self.log_dir if self.log_dir is not None else Path(
                paras.log_root, LOG_DIR, self.train_type, s['setting'], paras.algo, paras.pretrain_suffix,
                self.tgt_accent, str(paras.runs))
            self.resume_model_path = root.joinpath('snapshot.latest')
            self.optimizer_path = root.joinpath('optimizer.latest')
            info_dict_path = root.joinpath('info_dict.latest')
            assert self.resume_model_path.exists(), f"{self.resume_model_path} not exists..."
            assert info_dict_path.exists(), f"PreTraining info {info_dict_path} not exists..."
            with open(root.joinpath('global_step')) as f:
                self.global_step = int(f.read().strip())
            with open(info_dict_path, 'rb') as fin:
                self.train_info = pickle.load(fin)
        self.dashboard = _NullDashboard()
        self.data_container = None
        self.data_dirs = ([Path(s['data_root'], a) for a in self.accents] if 'data_root' in s else [])

    def load_data(self):
        """pretrain_interface.py:110-126: one endless bucketed train iterator and one dev loader per accent
        (data.DataContainer keeps the reference's on-disk format and batch order); synthetic runs and tests inject
        `self.data_container` themselves and leave `solver.data_root` out."""
        self.id2ch = self.id2units
        if not self.data_dirs:
            return
        from .data import DataContainer
        s = self.config['solver']
        self.data_container = DataContainer(self.data_dirs, batch_size=s['batch_size'], dev_batch_size=s['dev_batch_size'],
                                            is_memmap=getattr(self.paras, 'is_memmap', True),
                                            is_bucket=getattr(self.paras, 'is_bucket', True),
                                            num_workers=getattr(self.paras, 'njobs', 0), min_ilen=s.get('min_ilen'),
                                            max_ilen=s.get('max_ilen'), half_batch_ilen=s.get('half_batch_ilen'))

    def write_log(self, k, v):
        if self.log_dir is not None:
            with open(self.log_dir.joinpath(k), 'a') as fout:
                print(f"{self.global_step} {v}", file=fout)

    def log_msg(self, lr=None):
        pass

    def write_tr_logs(self):
        for k, v in self.train_info.items():
            self.write_log(f"train_{k}", float(v))

    def check_evaluate(self):
        if self.global_step % self.eval_ival == 0:
            self.evaluate()

    def write_dev_logs(self, prefix, info):
        for k, v in info.items():
            self.write_log(f"{prefix}_{k}", float(v))

    def save_best_model(self, tpe='wer', only_stat=False):
        """pretrain_interface / fo_meta_interface.py:56-68: model.<tpe>.best + the best_<tpe> stat file."""
        if self.log_dir is None:
            return
        if not only_stat:
            sd = OrderedDict((k, v.detach().cpu().clone()) for k, v in self.asr_model.state_dict().items())
            torch.save(sd, self.log_dir.joinpath(f'model.{tpe}.best'))
        with open(self.log_dir.joinpath(f'best_{tpe}'), 'w') as fout:
            print('{} {}'.format(self.global_step, getattr(self, f'best_{tpe}')), file=fout)

    def evaluate(self):
        """fo_meta_interface.py:253-298 / multi_interface.py:142-192: dev loop per pretraining accent with
        run_batch(train=False) under no_grad, per-accent and averaged logs, best WER / CER bookkeeping (when a scorer
        is configured: cer / wer are nan for synthetic vocabularies and nothing is ranked)."""
        self.write_tr_logs()
        dc = self.data_container
        if dc is None or not getattr(dc, 'dev_loaders', None) or not hasattr(self, '_eval'):
            return
        self.asr_model.eval()
        dev_info_ls = [RunningAvgDict(decay_rate=1.) for _ in range(self.num_pretrain)]
        for idx, dev_loader in enumerate(dc.dev_loaders):
            with torch.no_grad():
                for cur_b, (x, ilens, ys, olens) in enumerate(dev_loader):
                    if self.dev_max_ilen and int(ilens.max()) > self.dev_max_ilen:
                        continue
                    dev_info_ls[idx].add(self._eval(idx, x, ilens, ys, olens), len(ys))
            self.write_dev_logs(f"dev_{self.accents[idx]}", dev_info_ls[idx])
        dev_avg_info = RunningAvgDict(decay_rate=1.0)
        for dev_info in dev_info_ls:
            dev_avg_info.add({k: float(v) for k, v in dev_info.items()})
        self.write_dev_logs("dev_avg", dev_avg_info)
        cur_cer, cur_wer = float(dev_avg_info.get('cer', float('nan'))), float(dev_avg_info.get('wer', float('nan')))
        if cur_wer < self.best_wer:
            self.best_wer = cur_wer
            self.save_best_model()
        if cur_cer < self.best_cer:
            self.best_cer = cur_cer
            self.save_best_model('cer', only_stat=True)
        self.asr_model.train()
        return dev_avg_info

    def save_per_steps(self):
        """snapshot.latest / snapshot.step.N / info_dict.latest / global_step, rank 0 only
        (fo_meta_interface.py:70-88); what is saved is asr_model = last task's fast weights."""
        osd = self.optimizer_state()          # (a collective when the meta-Adam moments are sharded: before the rank filter)
        if self.log_dir is None:
            return
        sd = OrderedDict((k, v.detach().cpu().clone()) for k, v in self.asr_model.state_dict().items())
        torch.save(sd, self.log_dir.joinpath("snapshot.latest"))
        with open(self.log_dir.joinpath("info_dict.latest"), 'wb') as f:
            pickle.dump(self.train_info, f)
        with open(self.log_dir.joinpath("global_step"), 'w') as f:
            print(self.global_step, file=f)
        torch.save(sd, self.log_dir.joinpath(f"snapshot.step.{self.global_step}"))
        if osd is not None:
            torch.save(osd, self.log_dir.joinpath("optimizer.latest"))

    def optimizer_state(self):
        """What --resume needs beyond the fast weights of snapshot.latest: the optimizer of the outer loop."""
        opt = getattr(self, 'asr_opt', None)
        if isinstance(opt, FlatNoamAdam):
            return {k: (v.detach().cpu().clone() if torch.is_tensor(v) else v) for k, v in opt.state_dict().items()}
        return None


# ======================================================================================= FOMAML / Reptile
class FOMetaMixin:
    """Hot hooks of FOMetaASRInterface on flat arenas."""

    def _fo_init(self):
        paras, config = self.paras, self.config
        assert paras.meta_k is not None
        self.meta_k = paras.meta_k
        self.meta_batch_size = paras.meta_batch_size if paras.meta_batch_size is not None else self.num_pretrain
        self.asr_model, self.asr_opt = None, None
        self._train = partial(self.run_batch, train=True)
        self._eval = partial(self.run_batch, train=False)
        self._updates, self._counter = None, 0
        am = config['asr_model']
        opt = am['meta']['optimizer_opt']
        self.inner_lr = am['d_model'] ** (-0.5) * opt['k'] * (opt['warmup_steps'] ** (-0.5))   # :42-45
        self.reptile_outer = am.get('reptile_outer', 'adam')       # 'adam' (row R default) | 'interp'
        self.reptile_eps = float(am.get('reptile_eps', 1.0))

    # -- fo_meta_interface.py:90-111
    def load_model(self):
        eng = self.asr_model.engine
        if getattr(self.paras, 'resume', False):
            self.asr_model.load_state_dict(torch.load(self.resume_model_path))
        am = self.config['asr_model']
        if am['meta_opt_cls'] != 'noam':
            raise NotImplementedError("Should use noam optimizer in outer loop transformer learning")
        # the META weights: a detached flat clone.  The reference's clone breaks weight tying into two
        # tensors that receive identical gradients and identical Adam states, i.e. stay bit-identical
        # forever -> one copy is exact.
        self._original_flat = eng.params.clone()
        self._nvls = self._setup_nvls(eng)               # several NVSwitch-connected ranks: symmetric arenas + multicast
        if self._nvls is not None:
            self._nvls['theta'][:eng.layout.total].copy_(self._original_flat)
            self._original_flat = self._nvls['theta'][:eng.layout.total]
        self._original = OrderedDict()
        for name in eng.layout.shapes:
            if name == "pos_encoder.pe":
                self._original[name] = eng.pe
            else:
                src = "char_trans.weight" if (eng.cfg.tie and name == "pre_embed.weight") else name
                self._original[name] = eng.layout.view(self._original_flat, src)
        opt = am['meta']['optimizer_opt']
        self.meta_opt = _MetaNoamAdam(eng.be, self._original_flat, opt['k'], am['d_model'], opt['warmup_steps'])
        if getattr(self.paras, 'resume', False) and Path(getattr(self, 'optimizer_path', '')).is_file():
            osd = torch.load(self.optimizer_path)
            self._original_flat.copy_(osd['original'])             # the META weights (snapshot.latest = fast weights)
            st = self.meta_opt.state
            st.m.copy_(osd['m']); st.v.copy_(osd['v'])
            st.t, self.meta_opt.step_num = int(osd['t']), int(osd['step_num'])
        # update arena: [n params | 1 slot for the task counter] -> a single all-reduce carries both
        if self._nvls is not None:
            self._upd_flat = self._nvls['upd'][:eng.layout.total + 64]
            self._upd_flat.zero_()
        else:
            self._upd_flat = torch.zeros(eng.layout.total + 64, dtype=torch.float32, device=eng.device)
        self._gnorm = torch.zeros(1, dtype=torch.float64, device=eng.device)
        io = am.get('inner_optimizer_opt', {})
        if am.get('inner_optimizer_cls', 'SGD') != 'SGD':
            raise NotImplementedError("fused inner loop implements torch.optim.SGD (fometa-hkust.yaml:2-6)")
        self.asr_opt = FlatInnerSGD(eng, self.inner_lr, io.get('momentum', 0.0), io.get('nesterov', False))
        self._stats_ring = torch.zeros(max(self.num_pretrain, 1), 8, dtype=torch.float64, device=eng.device)
        self._ring_sizes = []

    def optimizer_state(self):
        """A collective when the Adam moments are sharded over the ranks (NVLS meta-update): every rank must call it."""
        st = self.meta_opt.state
        m, v = st.m, st.v
        nv = getattr(self, '_nvls', None)
        if nv is not None and nv['sharded']:
            import torch.distributed as tdist
            n, per, w, r = m.numel(), nv['per'], D.world_size(), D.rank()
            full = []
            for t in (m, v):
                mine = torch.zeros(per, dtype=t.dtype, device=t.device)
                a, b = r * per, min((r + 1) * per, n)
                if b > a:
                    mine[:b - a].copy_(t[a:b])
                allt = torch.empty(per * w, dtype=t.dtype, device=t.device)
                tdist.all_gather_into_tensor(allt, mine)
                full.append(allt[:n])
            m, v = full
        return {'original': self._original_flat.detach().cpu().clone(), 'm': m.detach().cpu().clone(),
                'v': v.detach().cpu().clone(), 't': st.t, 'step_num': self.meta_opt.step_num}

    def _setup_nvls(self, eng):
        """Symmetric (peer-mapped, NVSwitch-multicast) allocations for the meta weights and the update arena, so that the
        outer update -- sum over ranks, average, Adam, new weights to every rank -- is ONE kernel over the multicast
        mappings (masr_nvls_reduce_adam) instead of NCCL all-reduce + Adam + memset.  None (the NCCL path stays) unless
        there are several CUDA ranks under NCCL whose allocations get a multicast mapping; `asr_model.nvls_meta_update:
        false` switches it off."""
        if not (D.is_dist() and eng.device.type == 'cuda' and hasattr(eng.be, 'lib')):
            return None
        if not bool(self.config['asr_model'].get('nvls_meta_update', True)):
            return None
        import torch.distributed as tdist
        if tdist.get_backend() != 'nccl':
            return None
        # Two agreement points, so that a rank on which a step fails never leaves the others inside a collective: (1) every
        # rank could allocate the symmetric arenas, (2) every rank got a multicast mapping out of the rendezvous.
        def agreed(flag):
            t = torch.tensor([1.0 if flag else 0.0], device=eng.device)
            tdist.all_reduce(t, op=tdist.ReduceOp.MIN)
            return float(t) > 0
        w = D.world_size()
        n = eng.layout.total
        per = -(-(n + 64) // w)
        per = -(-per // 1024) * 1024                 # slice per rank, 4 KB granular
        theta = upd = None
        try:
            import torch.distributed._symmetric_memory as symm_mem
            theta = symm_mem.empty(per * w, dtype=torch.float32, device=eng.device)
            upd = symm_mem.empty(per * w, dtype=torch.float32, device=eng.device)
        except Exception as e:                                # no symmetric memory in this build / on this box: keep NCCL
            self._nvls_error = repr(e)
        if not agreed(theta is not None and upd is not None):
            return None
        ht = hu = None
        try:
            ht = symm_mem.rendezvous(theta, tdist.group.WORLD)
            hu = symm_mem.rendezvous(upd, tdist.group.WORLD)
        except Exception as e:
            self._nvls_error = repr(e)
        if not agreed(ht is not None and hu is not None and bool(ht.multicast_ptr) and bool(hu.multicast_ptr)):
            return None
        theta.zero_(); upd.zero_()
        return {'theta': theta, 'upd': upd, 'ht': ht, 'hu': hu, 'per': per, 'sharded': False}

    # -- lanes: independent accents of a meta-batch may run CONCURRENTLY on one GPU (asr_model.task_lanes > 1)
    def _lane(self, i):
        """Lane 0 is the solver's own engine / optimizer / accumulators; further lanes own a private engine
        (fast weights, gradients, workspaces, CUDA graphs, dropout seed), inner optimizer, norm and update
        arena, and a CUDA stream.  Tasks only share the read-only meta weights `_original_flat`."""
        lanes = self.__dict__.setdefault('_lanes', [])
        while len(lanes) <= i:
            k = len(lanes)
            eng0 = self.asr_model.engine
            if k == 0:
                lanes.append(_Lane(eng0, self.asr_opt, self._gnorm, self._upd_flat, None))
                continue
            from .engine import TransformerEngine
            eng = TransformerEngine(eng0.cfg, eng0.be, eng0.device, eng0.eps_ls, seed=531 + 7919 * k)
            eng.use_graphs, eng.multi_stream = eng0.use_graphs, eng0.multi_stream
            eng.seed_t.fill_(k << 40)
            io = self.config['asr_model'].get('inner_optimizer_opt', {})
            sgd = FlatInnerSGD(eng, self.inner_lr, io.get('momentum', 0.0), io.get('nesterov', False))
            lanes.append(_Lane(eng, sgd, torch.zeros_like(self._gnorm), torch.zeros_like(self._upd_flat),
                               torch.cuda.Stream(eng0.device)))
        return lanes[i]

    def _train_batch(self, lane, idx, x, ilens, ys, olens, accent_idx=None):
        """run_batch(train=True, sync=False) on a lane's engine."""
        if lane.stream is None:
            return self.run_batch(idx, x, ilens, ys, olens, train=True, accent_idx=accent_idx, sync=False)
        eng = lane.eng
        if isinstance(x, dict):
            db = x
        else:
            hb = eng.prepare_batch(x, ilens, ys, olens)
            db = hb if eng.use_graphs else eng.to_device(hb)
        eng.weights_dirty = True
        eng.forward_backward(db)

    # -- fo_meta_interface.py:223-250
    def run_task(self, batches, lane=None):
        lane = lane or self._lane(0)
        eng = lane.eng
        be = eng.be
        self._counter += 1
        eng.load_flat(self._original_flat)                     # load_state_dict(_original): one flat pass (+ bf16 shadow)
        if lane.stream is None:
            self.asr_model.train()
        eng.training = True
        lane.sgd.reset()                                       # fresh SGD per task
        for (idx, (x, ilens, ys, olens)) in batches:
            self._train_batch(lane, idx, x, ilens, ys, olens)
            be.mt_sumsq(eng.grads[:eng.layout.total], lane.gnorm)           # clip_grad_norm_ (device-side norm)
            lane.sgd.step(lane.gnorm, GRAD_CLIP)                             # NaN norm -> kernel skips the step

    # -- fo_meta_interface.py:145-154 (inner-loop test) + :180-198
    def inner_test(self, val_batch, lane=None, slot=None):
        lane = lane or self._lane(0)
        eng = lane.eng
        be = eng.be
        idx, (x, ilens, ys, olens) = val_batch
        if self.paras.algo == 'fomaml':
            self._train_batch(lane, idx, x, ilens, ys, olens, accent_idx=idx)
            be.mt_sumsq(eng.grads[:eng.layout.total], lane.gnorm)
        else:   # reptile: the held-out batch only produces logging statistics (no gradient needed)
            db = x if isinstance(x, dict) else eng.to_device(eng.prepare_batch(x, ilens, ys, olens))
            eng.weights_dirty = True
            eng.forward(db, want_grad=False)
        if slot is None:
            slot = len(self._ring_sizes)
            self._ring_sizes.append(len(ys))
        self._stats_ring[slot].copy_(eng.stats)
        self._partial_meta_update(lane)

    def _partial_meta_update(self, lane=None):
        lane = lane or self._lane(0)
        eng = lane.eng
        n = eng.layout.total
        self._updates = self._upd_flat
        if self.paras.algo == 'fomaml':       # _updates[n] += clip(p.grad)
            eng.be.mt_accumulate(lane.upd[:n], eng.grads[:n], lane.gnorm, GRAD_CLIP)
        elif self.paras.algo == 'reptile':    # _updates[n] += theta - phi
            eng.be.mt_reptile_delta(lane.upd[:n], self._original_flat[:n], eng.params[:n])
        else:
            raise ValueError(f"Not support meta algo {self.paras.algo}")

    # -- fo_meta_interface.py:200-221 (+ the one collective of the path)
    def _reduce_updates(self):
        """The one collective of the path: all-reduce(SUM) of the flat update arena (+ the task counter in its last
        slot).  Returns the number of tasks that contributed over all ranks."""
        eng = self.asr_model.engine
        n = eng.layout.total
        self._mark('reduce0')
        if D.is_dist():
            self._upd_flat[n] = float(self._counter)
            D.all_reduce_sum_(self._upd_flat)
            count = float(self._global_task_count) if self._global_task_count else float(self._upd_flat[n].item())
        else:
            count = float(self._counter)
        self._mark('reduce1')
        return count

    _phase_log = None            # list -> (tag, CUDA event) per phase boundary of a meta-step (bench.py: phase split)

    def _mark(self, tag):
        if self._phase_log is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self._phase_log.append((tag, ev))

    def _nvls_meta_update(self):
        """reduce + average + Adam + broadcast of the new meta weights + clearing of the update arena in one kernel."""
        eng, nv = self.asr_model.engine, self._nvls
        n, per, r = eng.layout.total, nv['per'], D.rank()
        opt = self.meta_opt
        opt.step_num += 1
        opt.lr = noam_lr(opt.step_num, opt.k, opt.d_model, opt.warmup_steps)
        st = opt.state
        st.t += 1
        bc1, bc2 = 1.0 - st.b1 ** st.t, 1.0 - st.b2 ** st.t
        stream = torch.cuda.current_stream(eng.device).cuda_stream
        self._mark('reduce0')
        nv['hu'].barrier(channel=0)               # every rank's update arena is final, nobody still reads the old weights
        rc = eng.be.lib.masr_nvls_reduce_adam(nv['hu'].multicast_ptr, nv['ht'].multicast_ptr, self._original_flat.data_ptr(),
                                              st.m.data_ptr(), st.v.data_ptr(), r * per, (r + 1) * per, n,
                                              float(max(self._global_task_count, 1)), float(opt.lr), st.b1, st.b2, st.eps,
                                              bc1, bc2, stream)
        if rc != 0:
            from . import _lib
            _lib.check(rc, "masr_nvls_reduce_adam")
        eng.be.launches += 1
        nv['hu'].barrier(channel=1)               # every slice of the new weights has landed everywhere
        nv['sharded'] = True
        self._mark('reduce1')

    def _final_meta_update(self):
        eng = self.asr_model.engine
        n = eng.layout.total
        if (getattr(self, '_nvls', None) is not None and self._global_task_count
                and not (self.paras.algo == 'reptile' and self.reptile_outer == 'interp')):
            self._nvls_meta_update()
            self._counter, self._updates = 0, None
            self._mark('adam1')
            return
        if getattr(self, '_nvls', None) is not None and self._nvls['sharded'] and self.paras.algo != 'reptile':
            raise RuntimeError("the meta-Adam moments are sharded over the ranks (NVLS meta-update): every meta-step needs "
                               "the global task count (meta_step_on_tasks(global_task_count=...))")
        count = self._reduce_updates()
        if self.paras.algo == 'reptile' and self.reptile_outer == 'interp':
            eng.be.mt_axpy(self._original_flat[:n], self._upd_flat[:n], -self.reptile_eps / max(count, 1.0))
            self.meta_opt.lr = self.reptile_eps
        else:
            self.meta_opt.step(self._upd_flat[:n], max(count, 1.0))
        eng.be.zero_(self._upd_flat)
        self._counter, self._updates = 0, None
        self._mark('adam1')

    _global_task_count = 0       # tasks of the meta-batch over ALL ranks (0: read it from the all-reduce)

    # -- host batches: staged on a copy stream, ahead of the compute that consumes them
    def stage_tasks(self, tasks, _prefetch=False):
        """Host batches of a meta-step -> device-resident prepared batches, copied on a private copy stream; every batch
        carries the event its consumer waits on, so the copy of the inner-test batch overlaps the inner-train batch and
        (through meta_step_on_tasks(next_tasks=)) the copies of step k+1 overlap step k.  Batches that are already
        prepared device dicts pass through."""
        eng = self.asr_model.engine
        if eng.device.type != 'cuda':
            return tasks
        if not any(not isinstance(b[1][0], dict) for tr, te in tasks for b in list(tr) + [te]):
            return tasks
        if '_copy_stream' not in self.__dict__:
            self._copy_stream = torch.cuda.Stream(eng.device)
            self._step_done = []               # end-of-step events of the last two meta-steps
        cs = self._copy_stream
        # staging blocks the allocator hands back may still be read by the kernels of the step that owned them: the copies
        # wait for the end of that step.  Staged at the head of a step: the previous step's buffers were just released.
        # Prefetched behind a step's launches: that step's buffers are still referenced, the ones released belong to the
        # step before it -- so these copies run under the compute of the step in flight.
        k = 2 if _prefetch else 1
        if len(self._step_done) >= k:
            cs.wait_event(self._step_done[-k])

        def stage(batch):
            idx, (x, ilens, ys, olens) = batch
            if isinstance(x, dict):
                return batch
            hb = eng.prepare_batch(x, ilens, ys, olens)
            db = eng.to_device(hb)
            ev = torch.cuda.Event()
            ev.record(cs)
            db["ready"] = ev
            return idx, (db, None, [None] * hb["B"], None)

        with torch.cuda.stream(cs):
            return [([stage(b) for b in tr], stage(te)) for tr, te in tasks]

    def meta_step_on_tasks(self, tasks, global_task_count=None, next_tasks=None):
        """One meta-step given this rank's tasks = [(train_batches, test_batch), ...]; batches are
        (accent_idx, (x, ilens, ys, olens)) tuples as DataContainer.get_item yields them.  next_tasks: the NEXT step's
        tasks, if the caller has them already (a prefetching loader): their host->device copies are issued behind this
        step's launches and run under its compute; pass the same list object as `tasks` of the next call."""
        pre = self.__dict__.pop('_prefetched', None)
        if pre is not None and pre[0] is tasks:
            tasks = pre[1]
        else:
            tasks = self.stage_tasks(tasks)
        self._meta_step_staged(tasks, global_task_count)
        if self.asr_model.engine.device.type == 'cuda' and '_copy_stream' in self.__dict__:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.asr_model.engine.device))
            self._step_done = (self._step_done + [ev])[-2:]
        if next_tasks is not None:
            self._prefetched = (next_tasks, self.stage_tasks(next_tasks, _prefetch=True))

    def _meta_step_staged(self, tasks, global_task_count=None):
        if global_task_count is not None:
            self._global_task_count = global_task_count
        self._ring_sizes = []
        n_lanes = min(int(self.config['asr_model'].get('task_lanes', 1)), len(tasks))
        be0 = self.asr_model.engine.be
        if hasattr(be0, 'gemm_stage_cap'):         # concurrent lanes share SMs: shallower operand rings (two CTAs per SM)
            be0.gemm_stage_cap(3 if n_lanes > 1 else 0)
        self.asr_model.engine.group_wgrads = False
        if n_lanes <= 1 or self.asr_model.engine.device.type != 'cuda':
            self._mark('step0')
            for tr_batches, val_batch in tasks:
                self.run_task(tr_batches)
                self._mark('train1')
                self.inner_test(val_batch)
                self._mark('test1')
        elif bool(self.config['asr_model'].get('lockstep', True)):
            self._meta_lockstep(tasks, n_lanes)
        else:
            # Accents are independent given the meta weights: run them on n_lanes CUDA streams so that the
            # small-kernel phases of one accent (decoder, LayerNorm, attention: a fraction of the 148 SMs)
            # overlap another accent's work.  The LAST task runs on lane 0, so asr_model still ends the
            # meta-step holding the last task's fast weights (fo_meta_interface.py:70-88, App. C #14).
            eng0 = self.asr_model.engine
            main = torch.cuda.current_stream(eng0.device)
            lanes = [self._lane(i) for i in range(n_lanes)]
            lane0_stream = torch.cuda.Stream(eng0.device) if not hasattr(self, '_lane0_stream') else self._lane0_stream
            self._lane0_stream = lane0_stream
            streams = [lane0_stream] + [l.stream for l in lanes[1:]]
            for st in streams:
                st.wait_stream(main)
            self._ring_sizes = [len(t[1][1][2]) for t in tasks]
            for i, (tr_batches, val_batch) in enumerate(tasks):
                li = (len(tasks) - 1 - i) % n_lanes
                with torch.cuda.stream(streams[li]):
                    self.run_task(tr_batches, lanes[li])
                    self.inner_test(val_batch, lanes[li], slot=i)
            for st in streams:
                main.wait_stream(st)
            n = eng0.layout.total
            for l in lanes[1:]:                      # combine the lanes' accumulators (then clear them)
                eng0.be.mt_axpy(self._upd_flat[:n], l.upd[:n], 1.0)
                eng0.be.zero_(l.upd)
        self._final_meta_update()

    def replica_checksum(self):
        """(sum, sum of squares) of the meta weights in float64, on the device: replicas of a multi-GPU run must agree
        bit for bit (every rank applies the identical average+Adam kernel to the identical all-reduced arena)."""
        w = self._original_flat.double()
        return torch.stack([w.sum(), (w * w).sum()])

    def _meta_lockstep(self, tasks, n_lanes):
        """Lock-step schedule of the tasks of a meta-step over n_lanes task lanes (private engines).

        A batch is three segments (engine.fb_segment): conv front end forward | encoder + decoder forward / backward |
        conv front end backward.  The conv segments are long launches that fill all 148 SMs; the middle segment is a
        dependent chain of ~150 launches of 36-150 CTAs each (latency bound).  Free-running lanes serialise: a
        persistent conv kernel of one lane leaves no SM for the other lanes' small kernels.  Here every round (the
        k-th inner-train batch of all lanes, then the inner-test batch) queues the conv segments of ALL lanes back to
        back on ONE stream and runs the middle segments CONCURRENTLY on the lanes' own streams, so the small kernels of
        different accents fill the machine together.  Same arithmetic per task as run_task / inner_test; the LAST task
        runs on lane 0 (asr_model ends the meta-step with its fast weights, fo_meta_interface.py:70-88)."""
        eng0 = self.asr_model.engine
        dev = eng0.device
        main = torch.cuda.current_stream(dev)
        if '_conv_stream' not in self.__dict__:
            self._conv_stream = torch.cuda.Stream(dev)
        if '_lane0_stream' not in self.__dict__:
            self._lane0_stream = torch.cuda.Stream(dev)
        conv = self._conv_stream
        lanes = [self._lane(i) for i in range(n_lanes)]
        for l in lanes:
            l.eng.group_wgrads = True          # several lanes share the GPU: one grouped weight-gradient launch per stretch
        streams = [self._lane0_stream] + [l.stream for l in lanes[1:]]
        for st in streams + [conv]:
            st.wait_stream(main)
        self._ring_sizes = [len(t[1][1][2]) for t in tasks]
        fomaml = self.paras.algo == 'fomaml'
        n = eng0.layout.total
        indexed = list(enumerate(tasks))
        for g0 in range(0, len(indexed), n_lanes):
            group = indexed[g0:g0 + n_lanes]
            assign = []
            for j, (slot, task) in enumerate(group):
                li = (len(group) - 1 - j) % n_lanes
                assign.append((lanes[li], streams[li], slot, task))
            for lane, st, slot, task in assign:                 # run_task prologue: load_state_dict(_original), fresh SGD
                with torch.cuda.stream(st):
                    self._counter += 1
                    lane.eng.load_flat(self._original_flat)
                    lane.eng.training = True
                    lane.sgd.reset()
            n_train = max(len(t[0]) for _, _, _, t in assign)
            for rnd in range(n_train + 1):
                test = rnd == n_train
                live = []
                for lane, st, slot, (tr, te) in assign:
                    if not test and rnd >= len(tr):
                        continue
                    idx, (x, ilens, ys, olens) = te if test else tr[rnd]
                    eng = lane.eng
                    with torch.cuda.stream(st):
                        if test and not fomaml:             # reptile: forward-only logging batch, then theta - phi
                            self.inner_test((idx, (x, ilens, ys, olens)), lane, slot=slot)
                            continue
                        if isinstance(x, dict):
                            db = x
                        else:
                            hb = eng.prepare_batch(x, ilens, ys, olens)
                            db = hb if eng.use_graphs else eng.to_device(hb)
                        h = eng.fb_begin(db)
                        ev = torch.cuda.Event()
                        ev.record(st)
                    live.append((lane, st, slot, h, ev))
                ev1s = []
                for lane, st, slot, h, ev in live:              # conv forward of every lane, back to back
                    conv.wait_event(ev)
                    with torch.cuda.stream(conv):
                        lane.eng.fb_segment(h, 0)
                        e1 = torch.cuda.Event()
                        e1.record(conv)
                    ev1s.append(e1)
                ev2s = []
                strict = bool(self.config['asr_model'].get('lockstep_strict', True))
                for (lane, st, slot, h, ev), e1 in zip(live, ev1s):   # the small-kernel chains, concurrently
                    # strict phases: no chain starts before the LAST conv segment is done -- a persistent conv kernel
                    # holds every SM, chains interleaved with it crawl one launch per conv kernel and delay its CTAs
                    st.wait_event(ev1s[-1] if strict else e1)
                    with torch.cuda.stream(st):
                        lane.eng.fb_segment(h, 1)
                        e2 = torch.cuda.Event()
                        e2.record(st)
                    ev2s.append(e2)
                if strict:
                    for e2 in ev2s:
                        conv.wait_event(e2)
                for (lane, st, slot, h, ev), e2 in zip(live, ev2s):   # conv backward, back to back again
                    conv.wait_event(e2)
                    with torch.cuda.stream(conv):
                        lane.eng.fb_segment(h, 2)
                        e3 = torch.cuda.Event()
                        e3.record(conv)
                    st.wait_event(e3)
                    eng = lane.eng
                    with torch.cuda.stream(st):
                        eng.be.mt_sumsq(eng.grads[:n], lane.gnorm)          # clip_grad_norm_ (device-side norm)
                        if test:
                            self._stats_ring[slot].copy_(eng.stats)
                            self._partial_meta_update(lane)
                        else:
                            # the task's last inner step: nothing reads its scaled gradient / momentum afterwards
                            lane.sgd.step(lane.gnorm, GRAD_CLIP, last=(rnd == n_train - 1))
        for st in streams + [conv]:
            main.wait_stream(st)
        for l in lanes[1:]:                      # combine the lanes' accumulators (then clear them)
            eng0.be.mt_axpy(self._upd_flat[:n], l.upd[:n], 1.0)
            eng0.be.zero_(l.upd)

    def flush_train_info(self):
        """The single device->host read of a meta-step: per-task inner-test loss/acc."""
        k = len(self._ring_sizes)
        infos = []
        if k:
            for row, bs in zip(self._stats_ring[:k].tolist(), self._ring_sizes):
                info = self.asr_model.engine.stats_to_info(row)
                infos.append(info)
                self.train_info.add(info, bs)
        self._ring_sizes = []
        return infos

    # -- fo_meta_interface.py:128-177
    def train(self):
        task_ids = list(range(self.num_pretrain))
        world = D.world_size()
        # One rank: the global python RNG, like the reference (fo_meta_interface.py:136), so the accent order of a
        # seeded run is the reference's.  Several ranks: every rank must draw the SAME permutation each step, but the
        # data pipeline (BucketSampler re-shuffles on loader reload) consumes the global RNG at rank-dependent
        # times -- sample tasks from a private stream that nothing else touches.
        task_rng = random if world == 1 else random.Random(int(getattr(self.paras, 'seed', 531)) * 7919 + 17)
        try:
            while self.global_step < self.max_step:
                for _ in range(self.eval_ival):
                    task_rng.shuffle(task_ids)         # identical on every rank
                    mine = D.partition_tasks(task_ids, self.meta_batch_size)
                    self._global_task_count = min(self.meta_batch_size, len(task_ids))
                    tasks = []
                    for accent_id in mine:
                        tr = self.data_container.get_item(accent_id, self.meta_k)
                        te = self.data_container.get_item(accent_id)[0]
                        tasks.append((tr, te))
                    self.meta_step_on_tasks(tasks)
                    self.flush_train_info()
                    self.log_msg(self.meta_opt.lr)
                    self.check_evaluate()
                    self.global_step += 1
                    self.dashboard.step()
                    if self.global_step % self.save_ival == 0:
                        self.save_per_steps()
                    self.dashboard.check()
                    if world > 1 and self.global_step >= self.max_step:
                        break
        except KeyboardInterrupt:
            self.save_per_steps()
            self.dashboard.set_status('pretrained(SIGINT)')
        else:
            self.dashboard.set_status('pretrained')


class _Lane:
    """One concurrent task slot of a meta-step (see FOMetaMixin._lane)."""

    def __init__(self, eng, sgd, gnorm, upd, stream):
        self.eng, self.sgd, self.gnorm, self.upd, self.stream = eng, sgd, gnorm, upd, stream


class _MetaNoamAdam:
    """meta_opt: noam schedule + Adam(0.9, 0.98, 1e-9) on the flat meta weights; averaging by the
    task count (`_updates /= _counter`) is fused into the Adam kernel."""

    def __init__(self, backend, original_flat, k, d_model, warmup_steps):
        self.k, self.d_model, self.warmup_steps = k, d_model, warmup_steps
        self.n = None
        self.state = FlatAdamState(backend, original_flat)
        self.step_num, self.lr = 0, d_model ** (-0.5)

    def step(self, upd_flat, count):
        self.step_num += 1
        self.lr = noam_lr(self.step_num, self.k, self.d_model, self.warmup_steps)
        n = upd_flat.numel()
        st = self.state
        st.t += 1
        bc1, bc2 = 1.0 - st.b1 ** st.t, 1.0 - st.b2 ** st.t
        st.be.mt_adam(st.p[:n], st.m[:n], st.v[:n], upd_flat, count, self.lr, st.b1, st.b2, st.eps, bc1, bc2)

    def zero_grad(self):
        pass


# ======================================================================================= multi-task
class MultiMixin:
    """Hot hooks of MultiASRInterface (multi_interface.py:94-140) on flat arenas."""

    def _multi_init(self):
        self.asr_model, self.asr_opt = None, None
        self._train = partial(self.run_batch, train=True)
        self._eval = partial(self.run_batch, train=False)
        if D.world_size() > 1:
            # data-parallel multi-task training: every rank must draw its OWN accents / batches (and dropout masks);
            # pretrain.py seeds all ranks identically, so re-seed the data-side RNG streams per rank
            import numpy as np
            seed = int(getattr(self.paras, 'seed', 531)) + 1000003 * D.rank()
            random.seed(seed); np.random.seed(seed % (2 ** 32)); torch.manual_seed(seed)

    def load_model(self):
        if getattr(self.paras, 'resume', False):
            self.asr_model.load_state_dict(torch.load(self.resume_model_path))
            if isinstance(self.asr_opt, FlatNoamAdam) and Path(getattr(self, 'optimizer_path', '')).is_file():
                self.asr_opt.load_state_dict(torch.load(self.optimizer_path))
        eng = self.asr_model.engine
        self._gnorm = torch.zeros(1, dtype=torch.float64, device=eng.device)
        if D.world_size() > 1:
            eng.seed_t.fill_(D.rank() << 40)          # rank-private dropout streams

    def multi_step(self, item, sync=True):
        """run_batch -> clip_grad_norm_ -> optimizer step (NaN norm skips the step on the device)."""
        eng = self.asr_model.engine
        idx, (x, ilens, ys, olens) = item
        info = self.run_batch(idx, x, ilens, ys, olens, train=True, accent_idx=idx, sync=sync)
        n = eng.layout.total
        world = D.world_size()
        if world > 1:                       # data-parallel variant: mean of the ranks' gradients
            D.all_reduce_sum_(eng.grads)
        eng.be.mt_sumsq(eng.grads[:n], self._gnorm)
        if isinstance(self.asr_opt, FlatNoamAdam):
            if world > 1:
                self.asr_opt.step_num += 1
                self.asr_opt.lr = noam_lr(self.asr_opt.step_num, self.asr_opt.k, self.asr_opt.d_model,
                                          self.asr_opt.warmup_steps)
                self.asr_opt.state.step(eng.grads, float(world), self.asr_opt.lr, skip_if_nan=self._gnorm,
                                        clip_sumsq=self._gnorm, max_norm=GRAD_CLIP * world)
                eng.weights_dirty = True
            else:
                self.asr_opt.step(self._gnorm, GRAD_CLIP)
        else:                               # any torch optimizer (fine-tuning configs): reference semantics
            eng.be.mt_clip(eng.grads[:n], self._gnorm, GRAD_CLIP)
            if not math.isnan(float(self._gnorm.item())):
                self.asr_opt.step()
            eng.weights_dirty = True
        return info

    def train(self):
        try:
            while self.global_step < self.max_step:
                for _ in range(self.eval_ival):
                    item = self.data_container.get_item()[0]
                    info = self.multi_step(item)
                    self.train_info.add(info, len(item[1][2]))
                    self.log_msg(getattr(self.asr_opt, 'lr', None))
                    self.check_evaluate()
                    self.global_step += 1
                    self.dashboard.step()
                    if self.global_step % self.save_ival == 0:
                        self.save_per_steps()
                    self.dashboard.check()
        except KeyboardInterrupt:
            self.save_per_steps()
            self.dashboard.set_status('pretrained(SIGINT)')
        else:
            self.dashboard.set_status('pretrained')


# ======================================================================================= fine-tune / mono
class TrainHost:
    """Minimal stand-alone equivalent of TrainInterface (src/train_interface.py:15-175): config fields, vocabulary,
    pretrain-model path, log-dir layout `LOG_DIR/<train_type>/<setting>/<algo>/<pretrain_suffix>/<eval_suffix>/<accent>/<runs>`,
    resume files (`snapshot.latest`, `optimizer.latest`, `info_dict.latest`, `epoch`, `global_step`, `best_{cer,wer}`).
    `self.train_set` / `self.dev_set` are iterables of (x, ilens, ys, olens) (io/dataset.py get_loader)."""

    def __init__(self, config, paras, id2accent):
        self.config, self.paras = config, paras
        self.train_type = 'evaluation'
        s = config['solver']
        self.eval_ival, self.log_ival = s['eval_ival'], s['log_ival']
        self.half_batch_ilen, self.dev_max_ilen = s.get('half_batch_ilen'), s.get('dev_max_ilen', 1 << 30)
        self.best_cer = self.best_wer = INIT_BEST_ER
        units = [SOS_SYMBOL]
        mapping = Path(s.get('spm_mapping', ''))
        if mapping.is_file():
            with open(mapping) as fin:
                units += [line.rstrip().split(' ')[0] for line in fin.readlines()]
        else:
            units += [f"<u{i}>" for i in range(1, int(s.get('n_units', 365)) + 1)]
        units.append(EOS_SYMBOL)
        self.id2units = self.id2ch = units
        self.metric_observer = _make_metric(s, units)
        self.save_verbose = getattr(paras, 'save_verbose', False)
        root = Path(getattr(paras, 'log_root', None) or Path.cwd())
        if getattr(paras, 'pretrain', False):
            if getattr(paras, 'pretrain_model_path', None):
                self.pretrain_model_path = Path(paras.pretrain_model_path)
            else:
                self.pretrain_model_path = Path(root, LOG_DIR, 'pretrain', paras.pretrain_setting, paras.algo,
                                                paras.pretrain_suffix, id2accent[paras.pretrain_tgt_accent],
                                                str(paras.pretrain_runs), f"snapshot.step.{paras.pretrain_step}")
            assert self.pretrain_model_path.exists(), f"Pretrain model path {self.pretrain_model_path} not exists"
            self.pretrain_module = s['pretrain_module']
        self.accent = id2accent[paras.accent]
        self.log_dir = None
        self.train_info = RunningAvgDict(decay_rate=0.99)
        self.global_step, self.ep = 1, 0
        if getattr(paras, 'log_root', None) is not None and D.rank() == 0:
            self.log_dir = Path(root, LOG_DIR, self.train_type, s['setting'], paras.algo, str(paras.pretrain_suffix),
                                str(getattr(paras, 'eval_suffix', None)), self.accent, str(paras.runs))
            if getattr(paras, 'resume', False):
                self.resume_model_path = self.log_dir.joinpath('snapshot.latest')
                self.optimizer_path = self.log_dir.joinpath('optimizer.latest')
                assert self.optimizer_path.exists(), f"Optimizer state {self.optimizer_path} not exists..."
                assert self.resume_model_path.exists(), f"{self.resume_model_path} not exists..."
                self.ep = int(self.log_dir.joinpath('epoch').read_text().strip())
                self.global_step = int(self.log_dir.joinpath('global_step').read_text().strip())
                for t in ('wer', 'cer'):
                    f = self.log_dir.joinpath(f'best_{t}')
                    if f.exists():
                        setattr(self, f'best_{t}', float(f.read_text().strip().split(' ')[1]))
                with open(self.log_dir.joinpath('info_dict.latest'), 'rb') as fin:
                    self.train_info = pickle.load(fin)
            else:
                self.log_dir.mkdir(parents=True, exist_ok=True)
        self.dashboard = _NullDashboard()
        self.train_set, self.dev_set = [], []
        self.data_dir = Path(s['data_root'], self.accent) if 'data_root' in s else None

    def load_data(self):
        """train_interface.py:129-154: bucketed train loader and sequential dev loader of the target accent."""
        self.id2ch = self.id2units
        if self.data_dir is None:
            return
        from .data import get_loader
        s = self.config['solver']
        memmap, njobs = getattr(self.paras, 'is_memmap', True), getattr(self.paras, 'njobs', 0)
        self.train_set = get_loader(self.data_dir.joinpath('train'), batch_size=s['batch_size'], min_ilen=s.get('min_ilen'),
                                    max_ilen=s.get('max_ilen'), half_batch_ilen=s.get('half_batch_ilen'),
                                    bucket_reverse=False, is_memmap=memmap,
                                    is_bucket=getattr(self.paras, 'is_bucket', True), num_workers=njobs)
        self.dev_set = get_loader(self.data_dir.joinpath('dev'), batch_size=s['dev_batch_size'], is_memmap=memmap,
                                  is_bucket=False, shuffle=False, num_workers=njobs)

    def write_log(self, k, v):
        if self.log_dir is not None:
            with open(self.log_dir.joinpath(k), 'a') as fout:
                print(f"{self.global_step} {v}", file=fout)

    def log_msg(self, lr=None):
        pass

    def write_logs(self, dev_info):
        for k, v in dev_info.items():
            self.write_log(f"dev_{k}", float(v))


class MonoMixin:
    """Hot hooks of MonoASRInterface (src/mono_interface.py:18-178): the fine-tune / mono-accent loop that follows
    pretraining in every experiment script.  One step = run_batch -> clip_grad_norm_(5) -> optimizer step on the flat
    arenas (same kernels as the multi-task step); `filter_model` / `freeze_module` / the per-epoch checkpoint keep the
    reference's names, files and state-dict keys."""

    def _mono_init(self):
        self.asr_model, self.asr_opt, self.lr_scheduler = None, None, None
        self.eval_every_epoch = getattr(self.paras, 'eval_every_epoch', False)
        self.max_epoch = self.config['solver'].get('total_epochs', 0)
        self._train = partial(self.run_batch, train=True)
        self._eval = partial(self.run_batch, train=False)
        self._frozen = []                    # (offset, numel) arena segments of frozen modules

    # ---- checkpoints (mono_interface.py:34-73)
    def save_per_epoch(self):
        if self.log_dir is None:
            return
        sd = OrderedDict((k, v.detach().cpu().clone()) for k, v in self.asr_model.state_dict().items())
        if self.save_verbose:
            torch.save(sd, self.log_dir.joinpath(f"snapshot.ep.{self.ep}"))
        torch.save(sd, self.log_dir.joinpath("snapshot.latest"))
        # the reference pickles its TransformerOptimizer object; here the optimizer is a flat-arena object whose
        # state_dict (step, lr, Adam m / v as CPU tensors) is what can be restored
        osd = self.asr_opt.state_dict()
        if isinstance(self.asr_opt, FlatNoamAdam):
            osd = {k: (v.detach().cpu().clone() if torch.is_tensor(v) else v) for k, v in osd.items()}
            with open(self.log_dir.joinpath("optimizer.latest"), "wb") as fout:
                pickle.dump(osd, fout)
        else:
            torch.save(osd, self.log_dir.joinpath("optimizer.latest"))
        with open(self.log_dir.joinpath("info_dict.latest"), 'wb') as fout:
            pickle.dump(self.train_info, fout)
        with open(self.log_dir.joinpath("global_step"), 'w') as fout:
            print(self.global_step, file=fout)
        with open(self.log_dir.joinpath("epoch"), 'w') as fout:
            print(self.ep, file=fout)

    def save_best_model(self, tpe='wer', only_stat=False):
        if self.log_dir is None:
            return
        if not only_stat:
            sd = OrderedDict((k, v.detach().cpu().clone()) for k, v in self.asr_model.state_dict().items())
            torch.save(sd, self.log_dir.joinpath(f'model.{tpe}.best'))
        with open(self.log_dir.joinpath(f'best_{tpe}'), 'w') as fout:
            print('{} {}'.format(self.global_step, getattr(self, f'best_{tpe}')), file=fout)

    def save_init(self):
        if self.log_dir is not None:
            torch.save(OrderedDict((k, v.detach().cpu().clone()) for k, v in self.asr_model.state_dict().items()),
                       self.log_dir.joinpath("snapshot.init"))

    # ---- model loading (mono_interface.py:75-116)
    def filter_model(self, state_dict):
        ret = OrderedDict()
        for k, v in state_dict.items():
            if k.split('.')[0] in self.pretrain_module:
                ret[k] = v
        return ret

    def load_model(self):
        eng = self.asr_model.engine
        self._gnorm = torch.zeros(1, dtype=torch.float64, device=eng.device)
        if getattr(self.paras, 'resume', False):
            self.asr_model.load_state_dict(torch.load(self.resume_model_path))
            if isinstance(self.asr_opt, FlatNoamAdam):
                with open(self.optimizer_path, 'rb') as fin:
                    self.asr_opt.load_state_dict(pickle.load(fin))
            else:
                self.asr_opt.load_state_dict(torch.load(self.optimizer_path))
        elif getattr(self.paras, 'pretrain', False):
            model_dict = OrderedDict((k, v.detach().clone()) for k, v in self.asr_model.state_dict().items())
            model_dict.update(self.filter_model(torch.load(self.pretrain_model_path)))
            self.asr_model.load_state_dict(model_dict)
            if 'freeze_module' in self.config['solver']:
                self.freeze_module(self.config['solver']['freeze_module'])

    def freeze_module(self, modules):
        """`p.requires_grad = False` for the module's parameters (mono_interface.py:109-116).  On the flat arenas a
        frozen tensor is a gradient segment that is zeroed before the norm / the optimizer: with Adam moments at 0
        the update of that segment is exactly 0, i.e. the parameter is skipped like in the reference."""
        eng = self.asr_model.engine
        for module in modules:
            for p in getattr(self.asr_model, module).parameters():
                p.requires_grad = False
            names = [n for n in eng.layout.offsets if n.split('.')[0] == module]
            if module == 'pre_embed' and 'pre_embed.weight' not in eng.layout.offsets:
                names.append('char_trans.weight')     # tied matrix: ONE Parameter in the reference, one arena tensor here
            for name in names:
                self._frozen.append((eng.layout.offsets[name], eng.layout.view(eng.grads, name).numel()))

    # ---- one step (mono_interface.py:131-148)
    def mono_step(self, cur_b, x, ilens, ys, olens, sync=True):
        eng = self.asr_model.engine
        info = self.run_batch(cur_b, x, ilens, ys, olens, train=True, sync=sync)
        for off, n in self._frozen:
            eng.grads[off:off + n].zero_()
        n = eng.layout.total
        eng.be.mt_sumsq(eng.grads[:n], self._gnorm)
        if isinstance(self.asr_opt, FlatNoamAdam):
            self.asr_opt.step(self._gnorm, GRAD_CLIP)       # NaN norm skips the step on the device
        else:                                               # any torch optimizer: reference semantics
            eng.be.mt_clip(eng.grads[:n], self._gnorm, GRAD_CLIP)
            if not math.isnan(float(self._gnorm.item())):
                self.asr_opt.step()
            eng.weights_dirty = True
        return info

    def check_evaluate(self):
        if self.global_step % self.eval_ival == 0:
            self.asr_opt.zero_grad()
            self.evaluate()

    def train(self):
        self.evaluate()
        try:
            if self.save_verbose:
                self.save_init()
            while self.ep < self.max_epoch:
                for cur_b, (x, ilens, ys, olens) in enumerate(self.train_set):
                    info = self.mono_step(cur_b, x, ilens, ys, olens)
                    self.train_info.add(info, len(ys))
                    self.log_msg(getattr(self.asr_opt, 'lr', None))
                    self.check_evaluate()
                    self.global_step += 1
                    self.dashboard.step()
                self.ep += 1
                self.save_per_epoch()
                self.dashboard.check()
                if self.eval_every_epoch:
                    self.evaluate()
        except KeyboardInterrupt:
            self.evaluate()
            self.dashboard.set_status('trained(SIGINT)')
        else:
            self.dashboard.set_status('trained')

    def evaluate(self):
        """Dev-set loop of mono_interface.py:180-230: run_batch(train=False) under no_grad, running averages, best
        CER / WER bookkeeping when a scorer (metric.Metric) is configured."""
        self.asr_model.eval()
        dev_info = RunningAvgDict(decay_rate=1.)
        with torch.no_grad():
            for cur_b, (x, ilens, ys, olens) in enumerate(self.dev_set):
                if int(ilens.max()) > self.dev_max_ilen:
                    continue
                dev_info.add(self._eval(cur_b, x, ilens, ys, olens), len(ys))
        self.write_logs(dev_info)
        for t in ('cer', 'wer'):
            if t in dev_info and float(dev_info[t]) < getattr(self, f'best_{t}'):
                setattr(self, f'best_{t}', float(dev_info[t]))
                self.save_best_model(t)
        self.asr_model.train()
        return dev_info


def fused(mixin, base):
    """Graft the fused hot hooks onto an interface base class (the reference's own
    FOMetaASRInterface / MultiASRInterface / MonoASRInterface, or PretrainHost / TrainHost)."""
    init_name = {FOMetaMixin: '_fo_init', MultiMixin: '_multi_init', MonoMixin: '_mono_init'}[mixin]

    class Fused(mixin, base):
        def __init__(self, config, paras, id2accent):
            base.__init__(self, config, paras, id2accent)
            getattr(self, init_name)()

    Fused.__name__ = f"Fused{base.__name__}"
    return Fused


FOMetaASRInterface = fused(FOMetaMixin, PretrainHost)
MultiASRInterface = fused(MultiMixin, PretrainHost)
MonoASRInterface = fused(MonoMixin, TrainHost)
