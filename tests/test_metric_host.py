"""Eval-path scoring (metaasr_crossaccent_b200/metric.py) against the reference's own Metric
(tests/golden/metric.npz, made with its sentencepiece model by oracle/make_golden.py golden_metric)."""
import numpy as np
import torch

from metaasr_crossaccent_b200.metric import Metric, levenshtein
from tests.helpers import GOLD


def test_levenshtein_known_answers():
    assert levenshtein("kitten", "sitting") == 3 and levenshtein("", "abc") == 3 and levenshtein("abc", "abc") == 0
    assert levenshtein(["a", "b", "c"], ["a", "c"]) == 1 and levenshtein([], []) == 0
    assert levenshtein("flaw", "lawn") == 2 and levenshtein("saturday", "sunday") == 3


def test_att_and_ctc_error_rates_match_reference():
    z = np.load(GOLD / "metric.npz")
    units = [str(u) for u in z["units"]]
    m = Metric(None, units, 0, len(units) - 1)            # no model file on this box: unigram decoding rule
    logits, ys = torch.from_numpy(z["att.logits"]), torch.from_numpy(z["att.ys"])
    assert m.batch_cal_er(logits, ys, ['att'], ['cer'])['att_cer'] == float(z["att.cer"])
    assert m.batch_cal_er(logits, ys, ['att'], ['wer'])['att_wer'] == float(z["att.wer"])
    both = m.batch_er_from_ids(torch.from_numpy(z["att.pred_ids"]), ys)       # one pass, ids from the CE kernel
    assert both == {"att_cer": float(z["att.cer"]), "att_wer": float(z["att.wer"])}
    for b in range(ys.shape[0]):
        assert m.cal_att_cer(torch.from_numpy(z["att.pred_ids"][b]), ys[b]) == float(z["att.cer_each"][b])
        assert m.cal_att_wer(torch.from_numpy(z["att.pred_ids"][b]), ys[b]) == float(z["att.wer_each"][b])
    cunits = ['<blank>'] + units[1:]
    mc = Metric(None, cunits, len(cunits) - 1, len(cunits) - 1)
    refs = np.split(z["ctc.refs"], np.cumsum(z["ctc.ref_lens"])[:-1])
    for b, r in enumerate(refs):
        p = torch.from_numpy(z["ctc.pred_ids"][b])
        assert mc.cal_ctc_cer(p, torch.from_numpy(r)) == float(z["ctc.cer_each"][b])
        assert mc.cal_ctc_wer(p, torch.from_numpy(r)) == float(z["ctc.wer_each"][b])


def test_discard_after_eos_quirks():
    """monitor/metric.py:24-33: the first position is never an eos candidate, a hypothesis without eos is EMPTY."""
    m = Metric(None, ['<s>', 'a', 'b', '</s>'], 0, 3)
    assert m.discard_ch_after_eos([1, 2, 3, 1]) == [1, 2] and m.discard_ch_after_eos([3, 1, 3]) == [3, 1]
    assert m.discard_ch_after_eos([1, 2, 1]) == [] and m.discard_ch_after_eos([1]) == []


def test_run_batch_eval_scores_from_ce_kernel_argmax():
    """run_batch(train=False) (transformer_torch_trainer.py:94-97 contract: {'cer','wer','loss','acc'}): the rates
    computed from the argmax ids of the fused CE kernel equal batch_cal_er on the logits."""
    from tests.test_interfaces_host import make_solver
    from tests.helpers import load_batch
    z = np.load(GOLD / "run_batch_tiny.npz")
    s = make_solver("fomaml")
    units = ['<s>'] + [('▁' if i % 3 == 0 else '') + chr(0x61 + i % 26) + str(i % 7) for i in range(365)] + ['</s>']
    s.metric_observer = Metric(None, units, 0, len(units) - 1)
    x, ilens, ys, olens = load_batch(z, "in.")
    with torch.no_grad():
        info = s.run_batch(0, x, ilens, ys, olens, train=False)
    assert set(info) == {"cer", "wer", "loss", "acc"}
    assert abs(info["loss"] - float(z["loss"])) <= 1e-5 * abs(float(z["loss"]))
    # the reference's own logits / padded targets of this batch (golden of the live reference)
    ref = s.metric_observer.batch_cal_er(torch.from_numpy(z["logit"]), torch.from_numpy(z["gold"]), ['att'], ['cer', 'wer'])
    assert info["cer"] == ref["att_cer"] and info["wer"] == ref["att_wer"]
    assert info["cer"] > 0 and info["wer"] > 0


def test_trim_and_best_hyp_format_match_reference_tester(tmp_path):
    """decode.trim / write_hyp against src/tester.py:189-207,271-273 (tests/golden/decode.npz)."""
    from metaasr_crossaccent_b200.decode import trim, write_hyp
    z = np.load(GOLD / "decode.npz")
    hyps = [h.tolist() for h in np.split(z["hyp_flat"], np.cumsum(z["hyp_lens"])[:-1])]
    for mn in ("transformer", "blstm"):
        ref = [h.tolist() for h in np.split(z[f"{mn}.flat"], np.cumsum(z[f"{mn}.lens"])[:-1])]
        assert [trim(list(h), mn, 366) for h in hyps] == ref
    for i, h in enumerate(hyps[:40]):
        write_hyp(tmp_path, [1 + i, 2, 3], trim(list(h), "transformer", 366))
    assert (tmp_path / "best-hyp").read_bytes() == z["best_hyp_bytes"].tobytes()
