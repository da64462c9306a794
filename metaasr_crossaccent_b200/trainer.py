"""Drop-in for the reference's src/transformer_torch_trainer.py: same `get_trainer(cls, config,
paras, id2accent)` mixin-by-closure contract (:13-18,107), same `set_model / exec / run_batch /
probe_model` hooks, same YAML keys.  `cls` may be the reference's own interface class
(FOMetaASRInterface, MultiASRInterface, MonoASRInterface) or the fused ones of interfaces.py.
"""
from __future__ import annotations

import torch

from . import ops
from .model import B200Transformer
from .optim import FlatNoamAdam, TransformerOptimizer

GRAD_CLIP = 5          # src/marcos.py:7
IGNORE_ID = -1         # src/marcos.py:8


def _make_backend(config, paras):
    """Optional, reference-preserving knobs: asr_model.dtype ('fp32' default | 'bf16'),
    asr_model.gemm ('simt' | 'umma'); device from LOCAL_RANK (one process per GPU)."""
    import os
    am = config["asr_model"]
    dtype = {"fp32": torch.float32, "bf16": torch.bfloat16}[am.get("dtype", "fp32")]
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
    factory = getattr(paras, "backend_factory", None)        # tests inject the torch double here
    if factory is not None:
        return factory(dtype)
    torch.cuda.set_device(dev)
    be = ops.CudaBackend(dev, dtype, gemm=am.get("gemm", "simt"))
    be.strict_umma = bool(am.get("strict_tcgen05", False))     # raise instead of warn when a tcgen05 kernel cannot be used
    return be


def get_trainer(cls, config, paras, id2accent):
    class TransformerTrainer(cls):
        def __init__(self, config, paras, id2accent):
            super(TransformerTrainer, self).__init__(config, paras, id2accent)

        def set_model(self):
            backend = _make_backend(self.config, self.paras)
            self.backend = backend
            self.label_smooth_rate = self.config['solver']['label_smoothing']
            self.asr_model = B200Transformer(self.id2ch, self.config['asr_model'], backend, backend.device,
                                             self.label_smooth_rate, getattr(self.paras, 'seed', 531))
            am = self.config['asr_model']
            self.asr_model.engine.use_graphs = bool(am.get('cuda_graphs', False))   # optional knob, default off
            if 'inner_optimizer_cls' not in am:                      # multi or mono
                if am['optimizer_cls'] == 'noam':
                    self.asr_opt = FlatNoamAdam(self.asr_model.engine, am['optimizer_opt']['k'], am['d_model'],
                                                am['optimizer_opt']['warmup_steps'])
                elif am['optimizer_cls'] == 'RAdam':
                    import torch_optimizer as extra_optim
                    self.asr_opt = getattr(extra_optim, am['optimizer_cls'])(self.asr_model.parameters(),
                                                                             **am['optimizer_opt'])
                else:
                    self.asr_opt = getattr(torch.optim, am['optimizer_cls'])(self.asr_model.parameters(),
                                                                             **am['optimizer_opt'])
            self.sos_id = self.asr_model.sos_id
            self.eos_id = self.asr_model.eos_id
            super().load_model()

        def exec(self):
            self.train()

        def run_batch(self, cur_b, x, ilens, ys, olens, train, accent_idx=None, sync=True):
            """Reference contract (transformer_torch_trainer.py:59-99).  train=True leaves the
            gradients of the batch on `asr_model.parameters()` (after an internal zero) and returns
            {'loss','acc'}; sync=False (fused interfaces) skips the device->host read and returns None,
            the statistics staying in engine.stats."""
            eng = self.asr_model.engine
            if isinstance(x, dict):                  # an already prepared, device-resident batch
                hb = db = x
            else:
                hb = eng.prepare_batch(x, ilens, ys, olens)
                # graph replay copies the host batch straight into its static device buffers
                db = hb if (train and eng.use_graphs) else eng.to_device(hb)
            if train:
                eng.weights_dirty = True
                ws = eng.forward_backward(db)
                self.asr_model.attach_grads()
                if not sync:
                    return None
                info = eng.read_stats()
                if self.global_step % 500 == 0 and getattr(self, 'metric_observer', None) is not None:
                    pred = ws["logits"].view(hb["B"], hb["L1"], -1)
                    self.probe_model(pred, db["ys_out"], accent_idx)
                return info
            eng.weights_dirty = True
            was_training = eng.training
            eng.training = False
            ws = eng.forward(db, want_grad=False)
            eng.training = was_training
            info = eng.read_stats()
            mo = getattr(self, 'metric_observer', None)
            cer = wer = float('nan')
            if mo is not None:
                if hasattr(mo, 'batch_er_from_ids'):
                    # the fused CE kernel already wrote the argmax ids: one [B, L+1] device->host copy scores both rates
                    er = mo.batch_er_from_ids(ws["argmax"].view(hb["B"], hb["L1"]), hb["ys_out"])
                    cer, wer = er['att_cer'], er['att_wer']
                else:                                # the reference's own Metric object (monitor/metric.py:36-48)
                    pred = ws["logits"].view(hb["B"], hb["L1"], -1)
                    cer = mo.batch_cal_er(pred, db["ys_out"], ['att'], ['cer'])['att_cer']
                    wer = mo.batch_cal_er(pred, db["ys_out"], ['att'], ['wer'])['att_wer']
            return {'cer': cer, 'wer': wer, 'loss': info['loss'], 'acc': info['acc']}

        def probe_model(self, pred, ys_out, accent_idx):
            self.metric_observer.cal_att_cer(torch.argmax(pred[0], dim=-1), ys_out[0], show=True, show_decode=True)
            self.metric_observer.cal_att_wer(torch.argmax(pred[0], dim=-1), ys_out[0], show=True)

    return TransformerTrainer(config, paras, id2accent)
