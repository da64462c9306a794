"""Eval-path scoring: drop-in for `src/monitor/metric.py` (`Metric.batch_cal_er`, `cal_att_{cer,wer}`,
`cal_ctc_{cer,wer}`, `discard_ch_after_eos`) -- character / word error rates of the greedy (argmax) hypotheses of a
batch against the references, through the sentencepiece unit inventory.

Same numbers as the reference (pinned by tests/golden/metric.npz, produced by the reference's Metric with its own
`valid_train_en_unigram150.model`), different plumbing:

  * the trainer hands over the argmax ids that the fused CE kernel already produced: ONE device->host copy of
    [B, L+1] integers per batch (`batch_er_from_ids`) instead of an argmax over the logits plus 2·B `.tolist()`
    synchronisations (the reference scores CER and WER in two separate passes);
  * `editdistance` (absent from this image) is replaced by a row-vectorised Levenshtein;
  * sentencepiece is used when the model file is there; otherwise pieces are decoded by the unigram rule that
    `DecodePieces` implements for ordinary pieces (concatenate, U+2581 -> space, drop the dummy prefix), which is
    what lets the fixtures travel to a box without the reference's data directory.
"""
from __future__ import annotations

from itertools import groupby
from pathlib import Path

import numpy as np
import torch

BLANK_SYMBOL = '<blank>'        # src/marcos.py
IGNORE_ID = -1
_WS = '▁'


def levenshtein(a, b) -> int:
    """Edit distance of two sequences (strings or lists), one numpy row per element of `a`."""
    n, m = len(a), len(b)
    if n == 0 or m == 0:
        return n + m
    if isinstance(a, str):
        ai, bi = np.frombuffer(a.encode('utf-32-le'), dtype=np.uint32), np.frombuffer(b.encode('utf-32-le'), dtype=np.uint32)
    else:
        voc = {}
        ai = np.fromiter((voc.setdefault(x, len(voc)) for x in a), dtype=np.int64, count=n)
        bi = np.fromiter((voc.setdefault(x, len(voc)) for x in b), dtype=np.int64, count=m)
    prev = np.arange(m + 1, dtype=np.int64)
    idx = np.arange(m + 1, dtype=np.int64)
    for i in range(1, n + 1):
        sub = prev[:-1] + (bi != ai[i - 1])
        cur = np.empty(m + 1, dtype=np.int64)
        cur[0] = i
        cur[1:] = np.minimum(sub, prev[1:] + 1)
        # insertions: cur[j] = min_k<=j (cur[k] + (j - k))  ->  running minimum of (cur - idx), plus idx
        cur = np.minimum.accumulate(cur - idx) + idx
        prev = cur
    return int(prev[-1])


class Metric:
    def __init__(self, model_path, id2units, sos_id, eos_id, ignore_id=None):
        self.spm = None
        if model_path is not None and Path(model_path).is_file():
            import sentencepiece as spmlib
            self.spm = spmlib.SentencePieceProcessor()
            self.spm.Load(str(model_path))
        self.id2units = list(id2units)
        self.sos_id, self.eos_id = sos_id, eos_id
        self.blank_id = None if BLANK_SYMBOL not in self.id2units else self.id2units.index(BLANK_SYMBOL)
        self.ignore_id = ignore_id

    # ---- text
    def decode_pieces(self, pieces):
        if self.spm is not None:
            return self.spm.DecodePieces(list(pieces))
        # control symbols (<s>, </s>, <blank>) decode to nothing, <unk> to its surface " ⁇ "; the dummy prefix is
        # dropped only when the first piece carries it
        pieces = [p for p in pieces if not (p.startswith('<') and p.endswith('>') and len(p) > 2 and p != '<unk>')]
        text = ''.join(' ⁇ ' if p == '<unk>' else p.replace(_WS, ' ') for p in pieces)
        return text[1:] if (pieces and pieces[0].startswith(_WS)) else text

    def discard_ch_after_eos(self, ls):
        eos_pos = 0
        if len(ls) == 1:
            return []
        for pos in range(1, len(ls)):
            if ls[pos] == self.eos_id:
                eos_pos = pos
                break
        return ls[:eos_pos]

    def _att_texts(self, pred_ids, y_ids):
        hyp = [x for x in self.discard_ch_after_eos(list(pred_ids)) if x != self.sos_id]
        hyp_text = self.decode_pieces([self.id2units[x] for x in hyp])
        ref_text = self.decode_pieces([self.id2units[x] for x in y_ids if x != self.eos_id and x != IGNORE_ID])
        return hyp_text, ref_text

    def _ctc_texts(self, pred_ids, y_ids):
        assert self.blank_id is not None
        hyp = [x[0] for x in groupby(list(pred_ids))]
        hyp = [x for x in hyp if x != self.sos_id and x != self.eos_id and x != self.blank_id]
        return (self.decode_pieces([self.id2units[x] for x in hyp]),
                self.decode_pieces([self.id2units[x] for x in y_ids]))

    @staticmethod
    def _ids(t):
        return t.tolist() if torch.is_tensor(t) else list(t)

    # ---- the reference's per-utterance entry points (monitor/metric.py:50-131)
    def cal_att_wer(self, pred, y, show=False, show_decode=False):
        h, r = self._att_texts(self._ids(pred), self._ids(y))
        h, r = h.split(' '), r.split(' ')
        return float(levenshtein(h, r)) / len(r) * 100

    def cal_att_cer(self, pred, y, show=False, show_decode=False):
        h, r = self._att_texts(self._ids(pred), self._ids(y))
        return float(levenshtein(h, r)) / len(r) * 100

    def cal_ctc_wer(self, pred, y, show=False, show_decode=False):
        h, r = self._ctc_texts(self._ids(pred), self._ids(y))
        h, r = h.split(' '), r.split(' ')
        return float(levenshtein(h, r)) / len(r) * 100

    def cal_ctc_cer(self, pred, y, show=False, show_decode=False):
        h, r = self._ctc_texts(self._ids(pred), self._ids(y))
        return float(levenshtein(h, r)) / len(r) * 100

    def batch_cal_er(self, preds, ys, modes, er_modes):
        """monitor/metric.py:36-48: preds = logits [B, L, C]; argmax here, one D2H for the whole batch."""
        pred = torch.argmax(preds, dim=-1)
        return self.batch_er_from_ids(pred, ys, modes, er_modes)

    def batch_er_from_ids(self, pred_ids, ys, modes=('att',), er_modes=('cer', 'wer')):
        """Same result as batch_cal_er for every (mode, er_mode), from hypothesis ids [B, L] (device or host tensor,
        e.g. the argmax output of the fused CE kernel) and references ys (tensor [B, L] or list of id sequences)."""
        pred_rows = pred_ids.detach().cpu().tolist() if torch.is_tensor(pred_ids) else [list(p) for p in pred_ids]
        y_rows = ys.detach().cpu().tolist() if torch.is_tensor(ys) else [self._ids(y) for y in ys]
        ret = {}
        for mode in modes:
            texts = [(self._att_texts if mode == 'att' else self._ctc_texts)(h, y) for h, y in zip(pred_rows, y_rows)]
            for er_mode in er_modes:
                er = 0.0
                for h, r in texts:
                    if er_mode == 'wer':
                        h, r = h.split(' '), r.split(' ')
                    er += float(levenshtein(h, r)) / len(r) * 100
                ret[f"{mode}_{er_mode}"] = er / len(pred_rows)
        return ret
