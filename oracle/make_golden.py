"""TEST INFRASTRUCTURE -- generates tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference, via oracle/ref_harness.py) on CPU in the build container.

    python oracle/make_golden.py            # rewrites every fixture

The fixtures are what pins oracle/port.py (and through it the CUDA path) to the reference:
the reference itself ships no tests or golden vectors.  Each fixture
carries its own weights and inputs, so nothing here has to be regenerated on the GPU box.
"""
from __future__ import annotations

import math
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle import ref_harness as H  # noqa: E402

GOLD = ROOT / "tests" / "golden"
TINY = dict(d_model=32, nheads=4, d_inner=64, enc_layers=2, dec_layers=2)


def synth_batch(g, B, T_list, L_list, idim=83):
    """Synthetic batch in the collate_fn layout (src/io/dataset.py:21-33): zero padded,
    sorted by ilen descending, labels uniform in [1, 365]."""
    Tm = max(T_list)
    x = torch.zeros(B, Tm, idim)
    for b, t in enumerate(T_list):
        x[b, :t] = torch.randn(t, idim, generator=g)
    ilens = torch.tensor(T_list, dtype=torch.int64)
    ys = [torch.randint(1, 366, (l,), generator=g, dtype=torch.int64) for l in L_list]
    olens = torch.tensor(L_list, dtype=torch.int64)
    return x, ilens, ys, olens


def pack_batch(prefix, batch, out):
    x, ilens, ys, olens = batch
    out[prefix + "x"] = x.numpy()
    out[prefix + "ilens"] = ilens.numpy()
    out[prefix + "olens"] = olens.numpy()
    out[prefix + "ys_cat"] = torch.cat(ys).numpy()


def sd_to_np(prefix, sd, out, skip_pe=True):
    for n, t in sd.items():
        if skip_pe and n == "pos_encoder.pe":
            continue
        out[prefix + n] = t.detach().cpu().numpy().copy()


def summarize(prefix, sd, out, skip_pe=True, nsample=384):
    """Compact pin of a dict of tensors: a strided sample + float64 sum and L2 norm each."""
    for n, t in sd.items():
        if skip_pe and n == "pos_encoder.pe":
            continue
        a = t.detach().cpu().numpy().reshape(-1)
        stride = max(1, a.size // nsample)
        out[prefix + n + "#sample"] = a[::stride][:nsample].copy()
        out[prefix + n + "#sum"] = np.float64(a.astype(np.float64).sum())
        out[prefix + n + "#l2"] = np.float64(np.sqrt((a.astype(np.float64) ** 2).sum()))


def clone_batch(b):
    x, ilens, ys, olens = b
    return x.clone(), ilens.clone(), [y.clone() for y in ys], olens.clone()


def golden_run_batch():
    """One run_batch(train=True) on a ragged batch: logits, targets, masks, loss, acc, grads."""
    cfg = H.base_config(**TINY, dropout=0.0)
    solver = H.build_reference_solver(cfg, H.make_paras("fomaml"))
    from src.nets_utils import make_bool_pad_mask, generate_square_subsequent_mask
    g = torch.Generator().manual_seed(1)
    batch = synth_batch(g, 3, [37, 30, 22], [5, 3, 4])
    out = {}
    pack_batch("in.", batch, out)
    w = {}
    sd_to_np("", solver.asr_model.state_dict(), w)
    np.savez_compressed(GOLD / "weights_tiny.npz", **w)
    out["pe_head"] = solver.asr_model.state_dict()["pos_encoder.pe"][:64, 0].numpy().copy()

    m = solver.asr_model
    m.train()
    x, ilens, ys, olens = clone_batch(batch)
    # a throw-away optimizer: run_batch calls self.asr_opt.zero_grad() (trainer :91)
    solver.asr_opt = torch.optim.SGD(m.parameters(), lr=0.0)
    logit, gold = m(x, ilens, ys, olens.clone())
    out["logit"] = logit.detach().numpy().copy()
    out["gold"] = gold.numpy().copy()
    enc_lens = torch.floor(ilens.to(dtype=torch.float32) / 4).to(dtype=torch.int64)
    out["enc_lens"] = enc_lens.numpy()
    out["enc_pad_mask"] = make_bool_pad_mask(enc_lens).numpy()
    out["causal_mask"] = generate_square_subsequent_mask(gold.shape[1]).numpy()
    ol = olens.clone()
    info = solver.run_batch(0, x, ilens, ys, ol, train=True)
    out["olens_after"] = ol.numpy()
    out["loss"] = np.float64(info["loss"])
    out["acc"] = np.float64(info["acc"])
    summarize("g.", {n: p.grad for n, p in m.named_parameters()}, out)
    # greedy decode (f-1 row): ids [L, B]
    m.eval()
    with torch.no_grad():
        out["greedy"] = m.recog(x, ilens).numpy().copy()
    np.savez_compressed(GOLD / "run_batch_tiny.npz", **out)
    print("run_batch_tiny: loss", info["loss"], "acc", info["acc"])


def golden_fomaml():
    """Two FOMAML meta-steps, 2 accents, meta_k=2, through the reference's own
    run_task/_train/_partial_meta_update/_final_meta_update (SURVEY App. A recipe)."""
    from torch import nn
    cfg = H.base_config(**TINY, dropout=0.0, warmup_steps=4, k=0.02)
    solver = H.build_reference_solver(cfg, H.make_paras("fomaml", meta_k=2))
    g = torch.Generator().manual_seed(2)
    out = {"meta_k": 2, "n_accents": 2, "n_meta_steps": 2, "warmup_steps": 4, "k": 0.02}
    w = np.load(GOLD / "weights_tiny.npz")
    for n, t in solver._original.items():        # same seed, same net -> same init
        if n != "pos_encoder.pe":
            assert np.array_equal(w[n], t.detach().numpy()), n
    shapes = [([24, 24], [4, 2]), ([31, 31, 31], [3, 5, 2]), ([40], [6])]
    for step in range(2):
        for acc in range(2):
            tr = [synth_batch(g, len(s[0]), *s) for s in (shapes[(step + acc) % 3], shapes[(step + acc + 1) % 3])]
            te = synth_batch(g, 2, [28, 20], [4, 3])
            for j, b in enumerate(tr):
                pack_batch(f"s{step}.a{acc}.tr{j}.", b, out)
            pack_batch(f"s{step}.a{acc}.te.", te, out)
            solver.run_task([(acc, clone_batch(b)) for b in tr])
            info = solver._train(acc, *clone_batch(te), accent_idx=acc)
            gn = nn.utils.clip_grad_norm_(solver.asr_model.parameters(), 5)
            assert not math.isnan(gn)
            out[f"s{step}.a{acc}.te_loss"] = np.float64(info["loss"])
            out[f"s{step}.a{acc}.te_gnorm"] = np.float64(float(gn))
            solver._partial_meta_update()
        # capture the averaged meta-gradient the reference feeds to Adam
        cnt = solver._counter
        summarize(f"s{step}.mg.", {n: u / cnt for n, u in solver._updates.items()}, out)
        solver._final_meta_update()
        out[f"s{step}.lr"] = np.float64(solver.meta_opt.lr)
        summarize(f"s{step}.w.", solver._original, out)
    summarize("fast.", solver.asr_model.state_dict(), out)
    out["inner_lr"] = np.float64(solver.inner_lr)
    np.savez_compressed(GOLD / "fomaml_tiny.npz", **out)
    print("fomaml_tiny: lr", solver.meta_opt.lr, "inner_lr", solver.inner_lr)


def golden_multi():
    """Three multi-task steps (multi_interface.py:100-114): run_batch, clip, noam-Adam."""
    from torch import nn
    cfg = H.base_config(**TINY, dropout=0.0, warmup_steps=4, k=0.02, meta=False)
    solver = H.build_reference_solver(cfg, H.make_paras("multi"))
    g = torch.Generator().manual_seed(3)
    out = {"n_steps": 3, "warmup_steps": 4, "k": 0.02}
    w = np.load(GOLD / "weights_tiny.npz")
    for n, t in solver.asr_model.state_dict().items():
        if n != "pos_encoder.pe":
            assert np.array_equal(w[n], t.detach().numpy()), n
    for step in range(3):
        b = synth_batch(g, 3, [33, 33, 26], [4, 6, 2])
        pack_batch(f"s{step}.", b, out)
        info = solver._train(0, *clone_batch(b), accent_idx=0)
        gn = nn.utils.clip_grad_norm_(solver.asr_model.parameters(), 5)
        summarize(f"s{step}.g.", {n: p.grad for n, p in solver.asr_model.named_parameters()}, out)
        solver.asr_opt.step()
        out[f"s{step}.loss"] = np.float64(info["loss"])
        out[f"s{step}.gnorm"] = np.float64(float(gn))
        summarize(f"s{step}.w.", solver.asr_model.state_dict(), out)
    np.savez_compressed(GOLD / "multi_tiny.npz", **out)
    print("multi_tiny done")


def golden_mono_freeze():
    """Fine-tune loop body of MonoASRInterface (mono_interface.py:75-116,131-148) run with the reference's own code:
    `filter_model` of a 'pretrained' state dict over pretrain_module, `freeze_module(['encoder'])`, then three
    steps of run_batch -> clip_grad_norm_(5) -> noam-Adam on the reference's model / optimizer objects."""
    from types import SimpleNamespace
    from torch import nn
    cfg = H.base_config(**TINY, dropout=0.0, warmup_steps=4, k=0.02, meta=False)
    solver = H.build_reference_solver(cfg, H.make_paras("multi"))
    from src.mono_interface import MonoASRInterface
    g = torch.Generator().manual_seed(11)
    out = {"n_steps": 3, "warmup_steps": 4, "k": 0.02}
    # a "pretrained" snapshot: the init weights plus a seeded perturbation; only pretrain_module entries are taken
    pre = {n: (t.detach().clone() + 0.01 * torch.randn(t.shape, generator=g) if n != "pos_encoder.pe" else t.detach().clone())
           for n, t in solver.asr_model.state_dict().items()}
    pre["pre_embed.weight"] = pre["char_trans.weight"]          # tied in the checkpoint like in the model
    sd_to_np("pre.", pre, out)
    solver.pretrain_module = ["feat_extractor", "vgg2enc", "encoder"]
    model_dict = solver.asr_model.state_dict()
    model_dict.update(MonoASRInterface.filter_model(solver, pre))
    solver.asr_model.load_state_dict(model_dict)
    MonoASRInterface.freeze_module(solver, ["encoder"])
    out["pretrain_module"] = np.array(solver.pretrain_module)
    out["freeze_module"] = np.array(["encoder"])
    for step in range(3):
        b = synth_batch(g, 3, [33, 33, 26], [4, 6, 2])
        pack_batch(f"s{step}.", b, out)
        info = solver._train(0, *clone_batch(b), accent_idx=0)
        gn = nn.utils.clip_grad_norm_(solver.asr_model.parameters(), 5)
        solver.asr_opt.step()
        out[f"s{step}.loss"] = np.float64(info["loss"])
        out[f"s{step}.gnorm"] = np.float64(float(gn))
        out[f"s{step}.lr"] = np.float64(solver.asr_opt.lr)
        summarize(f"s{step}.w.", solver.asr_model.state_dict(), out)
        summarize(f"s{step}.g.", {n: (p.grad if p.grad is not None else torch.zeros_like(p))
                                  for n, p in solver.asr_model.named_parameters()}, out)
    np.savez_compressed(GOLD / "mono_freeze_tiny.npz", **out)
    print("mono_freeze_tiny done")



def golden_hkust():
    """BASELINE shape (hkust net d512/h8/ff2048/2e4d, B=32, T=512, L=32) through the LIVE reference, summaries only
    (weights and inputs are regenerated from seeds by the tests: port.init_state_dict(seed 7),
    tests.helpers.hkust_profile_batch): one run_batch on the equal-length profile, one on the ragged profile, and
    one FOMAML meta-step (2 accents, meta_k = 1; accent 0 equal-length, accent 1 ragged) through the reference's own
    run_task / _train / clip_grad_norm_ / _partial_meta_update / _final_meta_update with k = 0.2, warmup = 4 so that
    the inner SGD step (4.4e-3) and the meta Adam step (1.1e-3) are far from trivial."""
    from torch import nn
    from oracle import port
    from tests.helpers import HKUST_K, HKUST_SEED_W, HKUST_WARMUP, hkust_profile_batch
    cfg = H.base_config(dropout=0.0, warmup_steps=HKUST_WARMUP, k=HKUST_K)
    solver = H.build_reference_solver(cfg, H.make_paras("fomaml", meta_k=1))
    sd = port.init_state_dict(port.NetCfg(), seed=HKUST_SEED_W)
    solver.asr_model.load_state_dict(sd)
    solver.load_model()                      # re-clone _original / rebuild meta_opt over the loaded weights
    m = solver.asr_model
    out = {"k": HKUST_K, "warmup_steps": HKUST_WARMUP, "seed_w": HKUST_SEED_W}
    solver.asr_opt = torch.optim.SGD(m.parameters(), lr=0.0)
    for tag, seed in (("eq", 101), ("rag", 102)):
        x, ilens, ys, olens = hkust_profile_batch(seed, tag)
        m.train()
        logit, gold = m(x, ilens, ys, olens.clone())
        out[f"{tag}.argmax"] = logit.detach().argmax(-1).numpy().astype(np.int16)
        srt = logit.detach().sort(-1).values
        out[f"{tag}.margin"] = (srt[..., -1] - srt[..., -2]).numpy().astype(np.float32)
        out[f"{tag}.gold"] = gold.numpy().astype(np.int16)
        out[f"{tag}.logit_absmax"] = np.float64(logit.detach().abs().max())
        summarize(f"{tag}.logit.", {"all": logit.detach()}, out, nsample=2048)
        out[f"{tag}.enc_lens"] = torch.floor(ilens.to(dtype=torch.float32) / 4).to(dtype=torch.int64).numpy()
        ol = olens.clone()
        info = solver.run_batch(0, x, ilens, ys, ol, train=True)
        out[f"{tag}.olens_after"] = ol.numpy()
        out[f"{tag}.loss"] = np.float64(info["loss"])
        out[f"{tag}.acc"] = np.float64(info["acc"])
        summarize(f"{tag}.g.", {n: p.grad for n, p in m.named_parameters()}, out)
        print("hkust", tag, info)
    # ---- one FOMAML meta-step
    batches = {0: (hkust_profile_batch(201, "eq"), hkust_profile_batch(202, "eq")),
               1: (hkust_profile_batch(203, "rag"), hkust_profile_batch(204, "rag"))}
    for acc in range(2):
        tr, te = batches[acc]
        solver.run_task([(acc, clone_batch(tr))])
        info = solver._train(acc, *clone_batch(te), accent_idx=acc)
        gn = nn.utils.clip_grad_norm_(solver.asr_model.parameters(), 5)
        assert not math.isnan(gn)
        out[f"fo.a{acc}.te_loss"] = np.float64(info["loss"])
        out[f"fo.a{acc}.te_gnorm"] = np.float64(float(gn))
        solver._partial_meta_update()
    cnt = solver._counter
    summarize("fo.mg.", {n: u / cnt for n, u in solver._updates.items()}, out)
    solver._final_meta_update()
    out["fo.lr"] = np.float64(solver.meta_opt.lr)
    out["fo.inner_lr"] = np.float64(solver.inner_lr)
    summarize("fo.w.", solver._original, out)
    summarize("fo.fast.", solver.asr_model.state_dict(), out)
    np.savez_compressed(GOLD / "hkust_b32.npz", **out)
    print("hkust_b32 done: lr", solver.meta_opt.lr, "inner_lr", solver.inner_lr)

def golden_loader():
    """Batch sequences of the reference's own input pipeline (src/io/dataset.py: BucketSampler + collate_fn +
    DataLoader, num_workers=0, and DataContainer.get_item) on synthetic accent directories that the test
    regenerates from the same seeds (tests/helpers.make_synth_accent_dir): per batch the dataset indices (marked in
    feat[first frame, 0]), ilens, olens and a checksum of the padded features."""
    import random
    import tempfile
    from pathlib import Path
    from tests.helpers import make_synth_accent_dir
    H.install_stubs()
    from src.io.dataset import DataContainer, get_loader
    out = {}
    with tempfile.TemporaryDirectory() as td:
        dirs = [make_synth_accent_dir(Path(td, f"acc{a}"), seed=100 + a) for a in range(2)]

        def record(prefix, batches):
            out[prefix + "n"] = np.int64(len(batches))
            for i, (x, ilens, ys, olens) in enumerate(batches):
                out[f"{prefix}{i}.idx"] = x[:, 0, 0].numpy().astype(np.int64)
                out[f"{prefix}{i}.ilens"] = ilens.numpy()
                out[f"{prefix}{i}.olens"] = olens.numpy()
                out[f"{prefix}{i}.xsum"] = np.float64(x.double().sum().item())
                out[f"{prefix}{i}.ysum"] = np.int64(sum(int(y.sum()) for y in ys))
                out[f"{prefix}{i}.shape"] = np.array(x.shape)
        # (1) bucketed train loader, two epochs, half batch beyond 50 frames, max_ilen cut
        random.seed(7); np.random.seed(7); torch.manual_seed(7)
        ld = get_loader(dirs[0] / "train", batch_size=8, is_memmap=True, is_bucket=True, num_workers=0, min_ilen=None,
                        max_ilen=70, half_batch_ilen=50)
        record("bucket.e0.", list(ld))
        record("bucket.e1.", list(ld))
        # (2) dev loader: sequential, no buckets
        record("dev.", list(get_loader(dirs[0] / "dev", batch_size=5, is_memmap=True, is_bucket=False, shuffle=False)))
        # (3) plain shuffled loader (torch RandomSampler)
        random.seed(9); np.random.seed(9); torch.manual_seed(9)
        record("shuf.", list(get_loader(dirs[1] / "train", batch_size=16, is_memmap=True, is_bucket=False, shuffle=True)))
        # (4) DataContainer.get_item: per-accent draws that run past the end of an epoch, then multi-task draws
        random.seed(11); np.random.seed(11); torch.manual_seed(11)
        dc = DataContainer(dirs, batch_size=8, dev_batch_size=5, is_memmap=True, is_bucket=True, num_workers=0,
                           max_ilen=70, half_batch_ilen=50)
        seq = [dc.get_item(accent_idx=a % 2, num=1)[0] for a in range(60)]
        seq += [it for _ in range(10) for it in dc.get_item(num=2)]
        out["dc.accents"] = np.array([a for a, _ in seq], dtype=np.int64)
        out["dc.reload_cnt"] = np.int64(dc.reload_cnt)
        record("dc.", [b for _, b in seq])
    np.savez_compressed(GOLD / "loader.npz", **out)
    print("loader done", len(out))


def golden_metric():
    """CER / WER of the reference's own Metric (src/monitor/metric.py, with its sentencepiece model and the repo's
    unigram150 unit list): attention-mode scores of seeded logits against padded references, CTC-mode scores of
    frame-level id sequences; the unit strings travel inside the fixture."""
    H.install_stubs()
    from src.monitor.metric import Metric
    data = H.REFERENCE_ROOT / "data"
    units = ['<s>'] + [l.rstrip().split(' ')[0] for l in open(data / "valid_train_en_unigram150_units.txt")] + ['</s>']
    m = Metric(str(data / "valid_train_en_unigram150.model"), units, 0, len(units) - 1)
    g = torch.Generator().manual_seed(17)
    B, L1, C = 6, 12, len(units)
    out = {"units": np.array(units)}
    # attention mode: hypotheses that partly agree with the references (so that the rates are not all > 100)
    ys = torch.full((B, L1), -1, dtype=torch.int64)
    for b in range(B):
        n = int(torch.randint(1, L1 - 1, (1,), generator=g))
        ys[b, :n] = torch.randint(2, C - 1, (n,), generator=g)
        ys[b, n] = C - 1
    logits = torch.randn(B, L1, C, generator=g)
    for b in range(B):
        for t in range(L1):
            if ys[b, t] >= 0 and float(torch.rand((), generator=g)) < 0.7:
                logits[b, t, ys[b, t]] += 8.0
    out["att.pred_ids"] = torch.argmax(logits, -1).numpy()
    out["att.ys"] = ys.numpy()
    out["att.logits"] = logits.numpy()
    out["att.cer"] = np.float64(m.batch_cal_er(logits, ys, ['att'], ['cer'])['att_cer'])
    out["att.wer"] = np.float64(m.batch_cal_er(logits, ys, ['att'], ['wer'])['att_wer'])
    out["att.cer_each"] = np.array([m.cal_att_cer(p, y) for p, y in zip(torch.argmax(logits, -1), ys)])
    out["att.wer_each"] = np.array([m.cal_att_wer(p, y) for p, y in zip(torch.argmax(logits, -1), ys)])
    # CTC mode (blstm vocabulary: <blank> first): frame-level ids with repeats and blanks
    cunits = ['<blank>'] + units[1:]
    mc = Metric(str(data / "valid_train_en_unigram150.model"), cunits, len(cunits) - 1, len(cunits) - 1)
    preds = torch.randint(0, C, (B, 30), generator=g)
    preds[:, ::3] = 0
    preds[:, 1::3] = preds[:, 2::3]
    refs = [torch.randint(2, C - 1, (int(torch.randint(1, 9, (1,), generator=g)),), generator=g) for _ in range(B)]
    out["ctc.pred_ids"] = preds.numpy()
    out["ctc.ref_lens"] = np.array([len(r) for r in refs])
    out["ctc.refs"] = torch.cat(refs).numpy()
    out["ctc.cer_each"] = np.array([mc.cal_ctc_cer(p, y) for p, y in zip(preds, refs)])
    out["ctc.wer_each"] = np.array([mc.cal_ctc_wer(p, y) for p, y in zip(preds, refs)])
    np.savez_compressed(GOLD / "metric.npz", **out)
    print("metric done", out["att.cer"], out["att.wer"])


def golden_decode():
    """Tester.trim / write_hyp (src/tester.py:189-207,271-273) on seeded hypotheses: the trimmed id lists for both
    model names and the bytes of the resulting `best-hyp` file."""
    import random
    import tempfile
    from pathlib import Path
    from types import SimpleNamespace
    H.install_stubs()
    from src.tester import Tester
    random.seed(3)
    hyps = [[random.choice([0, 5, 9, 366, 366, 12]) for _ in range(random.randint(0, 8))] for _ in range(200)]
    out = {"hyp_flat": np.array([x for h in hyps for x in h], dtype=np.int64), "hyp_lens": np.array([len(h) for h in hyps])}
    for mn in ("transformer", "blstm"):
        tr = [Tester.trim(SimpleNamespace(model_name=mn, eos_id=366), list(h)) for h in hyps]
        out[f"{mn}.flat"] = np.array([x for h in tr for x in h], dtype=np.int64)
        out[f"{mn}.lens"] = np.array([len(h) for h in tr])
    with tempfile.TemporaryDirectory() as td:
        for i, h in enumerate(hyps[:40]):
            Tester.write_hyp(SimpleNamespace(decode_dir=td), [1 + i, 2, 3], Tester.trim(SimpleNamespace(model_name="transformer", eos_id=366), list(h)))
        out["best_hyp_bytes"] = np.frombuffer(Path(td, "best-hyp").read_bytes(), dtype=np.uint8)
    np.savez_compressed(GOLD / "decode.npz", **out)
    print("decode done")


def golden_ctc():
    """The CTC call site of src/blstm_trainer.py:55-70 on synthetic encoder outputs:
    targets [366]+y+[366], blank 0, reduction='mean', zero_infinity=True.  Cases: ragged
    input lengths, repeated labels, an infeasible utterance (S > T), and L=0."""
    H.install_stubs()
    import torch.nn.functional as F
    ctc = torch.nn.CTCLoss(blank=0, reduction='mean', zero_infinity=True)
    g = torch.Generator().manual_seed(4)
    out = {}
    cases = {
        "a": dict(T=[20, 17, 12, 9], ys=[[5, 5, 7], [9], [3, 4, 4, 4, 2], [1, 2, 3, 4, 5, 6, 7, 8]], C=40),
        "b": dict(T=[33, 33], ys=[[11, 12, 13, 14, 15, 16], [2, 2, 2]], C=367),
    }
    for name, c in cases.items():
        B, Tm, C = len(c["T"]), max(c["T"]), c["C"]
        eos = C - 1
        logits = torch.randn(B, Tm, C, generator=g) * 2.0
        ys = [torch.tensor(y, dtype=torch.int64) for y in c["ys"]]
        ys_out = [torch.cat([torch.tensor([eos]), y, torch.tensor([eos])]) for y in ys]
        y_true = torch.cat(ys_out)
        olens = torch.tensor([len(y) for y in ys_out], dtype=torch.int64)
        enc_lens = torch.tensor(c["T"], dtype=torch.int64)
        lg = logits.clone().requires_grad_(True)
        pred = F.log_softmax(lg, dim=-1)
        loss = ctc(pred.transpose(0, 1).contiguous(), y_true, enc_lens, olens)
        loss.backward()
        nll = F.ctc_loss(pred.detach().transpose(0, 1).contiguous(), y_true, enc_lens, olens,
                         blank=0, reduction='none', zero_infinity=True)
        out[f"{name}.logits"] = logits.numpy()
        out[f"{name}.targets"] = y_true.numpy()
        out[f"{name}.in_lens"] = enc_lens.numpy()
        out[f"{name}.tgt_lens"] = olens.numpy()
        out[f"{name}.loss"] = np.float64(float(loss))
        out[f"{name}.nll"] = nll.numpy()
        out[f"{name}.grad_logits"] = lg.grad.numpy().copy()
        print(f"ctc case {name}: loss {float(loss):.6f} nll {nll.tolist()}")
    np.savez_compressed(GOLD / "ctc.npz", **out)


if __name__ == "__main__":
    assert H.reference_available(), "needs /root/reference (build container only)"
    GOLD.mkdir(parents=True, exist_ok=True)
    torch.set_num_threads(8)
    if len(sys.argv) > 1:                    # e.g. `python oracle/make_golden.py hkust`: one fixture only
        for nm in sys.argv[1:]:
            globals()["golden_" + nm]()
        sys.exit(0)
    golden_run_batch()
    golden_fomaml()
    golden_multi()
    golden_mono_freeze()
    golden_loader()
    golden_metric()
    golden_decode()
    golden_ctc()
    golden_hkust()
