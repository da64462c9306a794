"""Greedy-decode driver of the test phase: `Tester.trim / batch_greedy_decode / write_hyp`
(src/tester.py:189-239,271-273) on top of `B200Transformer.recog` (the CUDA path, encoder memory computed once).

The `best-hyp` file is the wire format consumed by translate.py:90-99: one line per utterance,
`"<ref ids separated by blanks>\\t<hyp ids separated by blanks>\\n"`, appended in batch order."""
from __future__ import annotations

from itertools import groupby
from pathlib import Path

import torch


def trim(hyp, model_name, eos_id):
    """src/tester.py:189-207.  transformer: cut at the first eos AFTER position 0 (position 0 is never tested; a
    one-element hypothesis is empty; without an eos everything is kept); blstm: drop ids >= eos."""
    assert isinstance(hyp, list)
    if model_name == 'blstm':
        return [i for i in hyp if i < eos_id]
    if model_name == 'transformer':
        eos_pos = len(hyp) + 1
        if len(hyp) > 1:
            for pos in range(1, len(hyp)):
                if hyp[pos] == eos_id:
                    eos_pos = pos
                    break
            return hyp[:eos_pos]
        return []
    raise NotImplementedError


def write_hyp(decode_dir, y, hyp):
    """src/tester.py:271-273."""
    with open(Path(decode_dir, 'best-hyp'), 'a') as fout:
        fout.write("{}\t{}\n".format(" ".join(str(i) for i in y), " ".join(str(i) for i in hyp)))


def batch_greedy_decode(asr_model, xs, ilens, ys, decode_dir, model_name='transformer', eos_id=None, blank_id=None):
    """src/tester.py:210-239.  transformer: ids = recog(xs, ilens) [L, B] -> per utterance trim + write; blstm: frame
    argmax -> trim -> collapse repeats -> drop blanks.  ONE device->host copy of the id matrix per batch."""
    eos_id = asr_model.eos_id if eos_id is None else eos_id
    with torch.no_grad():
        if model_name == 'transformer':
            preds = asr_model.recog(xs, ilens).transpose(0, 1).cpu().tolist()
            for pred, y in zip(preds, ys):
                write_hyp(decode_dir, y.tolist(), trim(pred, model_name, eos_id))
        elif model_name == 'blstm':
            out, _ = asr_model(xs, ilens)
            for pred, y in zip(torch.argmax(out, dim=-1).cpu().tolist(), ys):
                pred = [x[0] for x in groupby(trim(pred, model_name, eos_id))]
                if blank_id is not None:
                    pred = [x for x in pred if x != blank_id]
                write_hyp(decode_dir, y.tolist(), pred)
        else:
            raise NotImplementedError(f"{model_name} doesn't support greedy decode batchwise")
    return True
