"""GPU parity of kernel 1 (CTC alpha-beta) through the C ABI: against the committed goldens of the
reference call site (nn.CTCLoss as configured in src/blstm_trainer.py:22), against the oracle port,
against F.ctc_loss on seeded inputs incl. the BASELINE shape, and the domain's size-independent
properties (gradient rows sum to zero; zero gradient beyond the input length)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import port
from tests.helpers import GOLD

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    return torch.device("cuda", 0)


def test_ctc_golden_reference_call_site(dev):
    from metaasr_crossaccent_b200.ctc import ctc_fwd_bwd
    z = np.load(GOLD / "ctc.npz")
    for c in ("a", "b"):
        logits = torch.from_numpy(z[f"{c}.logits"].copy()).transpose(0, 1).contiguous().to(dev)   # [T,B,C]
        tg, il, tl = (torch.from_numpy(z[f"{c}.{k}"].copy()) for k in ("targets", "in_lens", "tgt_lens"))
        loss, nll, grad = ctc_fwd_bwd(logits, tg, il, tl, blank=0, zero_infinity=True, is_logprob=False)
        assert np.allclose(nll.cpu().numpy(), z[f"{c}.nll"], rtol=1e-5, atol=1e-4)
        ref = float(z[f"{c}.loss"])
        assert abs(float(loss) - ref) <= 1e-5 * abs(ref)
        g_ref = z[f"{c}.grad_logits"]
        g = grad.transpose(0, 1).cpu().numpy()
        assert np.abs(g - g_ref).max() <= 1e-4 * np.abs(g_ref).max()
        assert (g[g_ref == 0] == 0).all()                      # zero pattern exact
        # log-prob input mode == nn.CTCLoss contract
        lp = torch.log_softmax(logits, -1)
        loss2, nll2, grad2 = ctc_fwd_bwd(lp.contiguous(), tg, il, tl, is_logprob=True)
        assert abs(float(loss2) - ref) <= 1e-5 * abs(ref)
        assert np.abs(grad2.transpose(0, 1).cpu().numpy() - g_ref).max() <= 1e-4 * np.abs(g_ref).max()


@pytest.mark.parametrize("T,B,C,L", [(128, 32, 367, 32), (50, 5, 40, 70), (375, 4, 367, 100), (20, 3, 10, 0)])
def test_ctc_vs_torch_and_properties(dev, T, B, C, L):
    from metaasr_crossaccent_b200.ctc import B200CTCLoss, ctc_fwd_bwd
    g = torch.Generator().manual_seed(T * 1000 + L)
    logits = (torch.randn(T, B, C, generator=g) * 2).to(dev)
    in_lens = torch.tensor([T - (3 * b) % max(T // 2, 1) for b in range(B)], dtype=torch.int64)
    tgt_lens = torch.tensor([max(0, L - b) for b in range(B)], dtype=torch.int64)
    ys = [torch.randint(1, C - 1, (int(l),), generator=g) for l in tgt_lens]
    if L >= 3:
        ys[0][1] = ys[0][0]                                  # repeated label
    e = torch.tensor([C - 1])
    ys_out = [torch.cat([e, y, e]) for y in ys]              # [366]+y+[366] (blstm_trainer.py:56-58)
    targets = torch.cat(ys_out)
    tl = torch.tensor([len(y) for y in ys_out], dtype=torch.int64)
    # reference: the same nn.CTCLoss configuration evaluated in float64
    lg = logits.double().requires_grad_(True)
    lp = F.log_softmax(lg, -1)
    ref = F.ctc_loss(lp, targets.to(dev), in_lens, tl, blank=0, reduction='mean', zero_infinity=True)
    ref.backward()
    loss, nll, grad = ctc_fwd_bwd(logits, targets, in_lens, tl)
    assert abs(float(loss) - float(ref)) <= 1e-5 * abs(float(ref)) + 1e-7
    scale = float(lg.grad.abs().max())
    # alpha/beta are fp32 log-probabilities of magnitude ~|nll|: posteriors carry ~8 ulp(|nll|) relative error
    gtol = max(2e-4, 8 * 1.2e-7 * float(nll.abs().max()))
    assert float((grad - lg.grad).abs().max()) <= gtol * scale + 1e-9
    # properties: rows of d/dlogits sum to 0; nothing beyond the input length
    assert float(grad.sum(-1).abs().max()) <= 1e-5 * scale * C
    for b in range(B):
        assert float(grad[int(in_lens[b]):, b].abs().max() if int(in_lens[b]) < T else 0.0) == 0.0
    # autograd module drop-in for nn.CTCLoss(blank=0, reduction='mean', zero_infinity=True)
    lg2 = logits.clone().requires_grad_(True)
    l2 = B200CTCLoss()(F.log_softmax(lg2, -1), targets, in_lens, tl)
    l2.backward()
    assert float((lg2.grad - lg.grad).abs().max()) <= gtol * scale + 1e-9
    # small case also against the pure-python oracle in float64
    if T * B * max(L, 1) <= 50 * 5 * 70:
        onll, oloss, ograd = port.ctc_alpha_beta(logits.cpu(), targets, in_lens, tl)
        assert abs(float(loss) - float(oloss)) <= 1e-5 * abs(float(oloss)) + 1e-7
        assert float((grad.cpu().double() - ograd).abs().max()) <= 2e-4 * scale + 1e-9


def test_ctc_infeasible_and_large_workspace(dev):
    from metaasr_crossaccent_b200.ctc import ctc_fwd_bwd
    # S > T: infeasible -> nll 0 and zero gradient under zero_infinity
    logits = torch.randn(4, 2, 12, device=dev)
    targets = torch.tensor([1, 2, 3, 4, 5, 6, 7, 1], dtype=torch.int64)
    loss, nll, grad = ctc_fwd_bwd(logits, targets, torch.tensor([4, 4]), torch.tensor([7, 1]))
    assert float(nll[0]) == 0.0 and float(grad[:, 0].abs().max()) == 0.0 and float(nll[1]) > 0
    # long utterances: tables go to the global workspace (T'=750, L=150)
    T, B, C, L = 750, 3, 367, 150
    g = torch.Generator().manual_seed(5)
    l32 = torch.randn(T, B, C, generator=g).to(dev)
    lg = l32.double().requires_grad_(True)
    tg = torch.randint(1, C, (B * L,), generator=g)
    il, tl = torch.tensor([750, 700, 600]), torch.tensor([L] * B)
    ref = F.ctc_loss(F.log_softmax(lg, -1), tg.to(dev), il, tl, blank=0, reduction='mean', zero_infinity=True)
    ref.backward()
    loss, nll, grad = ctc_fwd_bwd(l32, tg, il, tl)
    assert abs(float(loss) - float(ref)) <= 1e-5 * abs(float(ref))
    gtol = max(2e-4, 8 * 1.2e-7 * float(nll.abs().max()))
    assert float((grad - lg.grad).abs().max()) <= gtol * float(lg.grad.abs().max())


@pytest.mark.parametrize("case", ["ragged_mixed", "many_repeats", "logprob_input"])
def test_ctc_long_utterances_frame_chunked_kernel(dev, case):
    """Utterances whose posterior table exceeds shared memory (T' = 375 .. 750, the reference's max_ilen 1500 / 3000) run
    on the frame-chunked kernel (alpha boundary rows in sweep A, chunk tables recomputed in sweep B): against float64
    F.ctc_loss, with in one batch a full-length utterance, odd chunk remainders, an utterance shorter than one chunk, a
    one-frame utterance, an empty target and an infeasible one; and targets with so many repeated classes that the
    side list of third-and-later occurrences overflows (per-frame search path)."""
    from metaasr_crossaccent_b200.ctc import ctc_fwd_bwd
    g = torch.Generator().manual_seed(11)
    C = 367
    if case == "many_repeats":
        T = 400
        ys = [torch.randint(1, 6, (150,), generator=g), torch.randint(1, 4, (120,), generator=g)]   # ~5 classes over 150 positions
        in_lens = torch.tensor([400, 391])
    else:
        T = 750 if case == "ragged_mixed" else 377
        lens = [152, 140, 3, 1, 0, 120, 90]
        ys = [torch.randint(1, C, (l,), generator=g) for l in lens]
        in_lens = torch.tensor([T, T - 33, 50, 1, 5, 100, T - 1])        # utterance 5: 120 labels in 100 frames -> infeasible
    B = len(ys)
    logits = (torch.randn(T, B, C, generator=g) * 2).to(dev)
    targets = torch.cat(ys)
    tl = torch.tensor([len(y) for y in ys], dtype=torch.int64)
    lg = logits.double().requires_grad_(True)
    lp = F.log_softmax(lg, -1)
    ref = F.ctc_loss(lp, targets.to(dev), in_lens, tl, blank=0, reduction='mean', zero_infinity=True)
    ref.backward()
    if case == "logprob_input":
        lp32 = F.log_softmax(logits, -1).contiguous()
        loss, nll, grad_lp = ctc_fwd_bwd(lp32, targets, in_lens, tl, is_logprob=True)
        # ATen's gradient w.r.t. the log-probabilities, pushed through log_softmax = the gradient w.r.t. the logits
        grad = grad_lp - torch.exp(lp32) * grad_lp.sum(-1, keepdim=True)
    else:
        loss, nll, grad = ctc_fwd_bwd(logits, targets, in_lens, tl)
    assert abs(float(loss) - float(ref)) <= 1e-5 * abs(float(ref))
    scale = float(lg.grad.abs().max())
    gtol = max(2e-4, 8 * 1.2e-7 * float(nll.abs().max()))
    assert float((grad - lg.grad).abs().max()) <= gtol * scale + 1e-9
    for b in range(B):
        if int(in_lens[b]) < T:
            assert float(grad[int(in_lens[b]):, b].abs().max()) == 0.0
    if case != "many_repeats":
        assert float(nll[5]) == 0.0 and float(grad[:, 5].abs().max()) == 0.0       # infeasible under zero_infinity
    # forward only (no gradient buffer) gives the same likelihoods
    loss2, nll2, _ = ctc_fwd_bwd(logits if case != "logprob_input" else lp32, targets, in_lens, tl,
                                 is_logprob=case == "logprob_input", want_grad=False)
    assert torch.equal(nll2, nll)


def _torch_ref(logits, targets, in_lens, tl, dev):
    lg = logits.double().requires_grad_(True)
    ref = F.ctc_loss(F.log_softmax(lg, -1), targets.to(dev), in_lens, tl, blank=0, reduction='mean', zero_infinity=True)
    ref.backward()
    return ref, lg.grad


@pytest.mark.parametrize("case", ["triple_repeats", "many_repeats", "blank_as_label", "tb1_and_short", "same_label_run"])
def test_ctc_v3_label_structure_edge_cases(dev, case):
    """The single-table kernel (ctc3.cu) treats a class's first / second occurrence in registers-free tables, later
    occurrences and labels equal to the blank class through a side list, and falls back to its generic gradient path
    when that list overflows: every one of those routes against F.ctc_loss in float64."""
    from metaasr_crossaccent_b200.ctc import ctc_fwd_bwd
    g = torch.Generator().manual_seed(11)
    C = 367
    if case == "triple_repeats":            # a few classes occur 3-5 times: side list
        T, ys = 96, [torch.tensor([5, 9, 5, 7, 5, 9, 9, 5, 11, 5, 12, 9]), torch.tensor([3, 3, 3, 3]), torch.tensor([8, 2, 8])]
    elif case == "many_repeats":            # 60 labels over 3 classes: the side list overflows -> generic path
        T, ys = 160, [torch.randint(1, 4, (60,), generator=g), torch.randint(1, 4, (50,), generator=g)]
    elif case == "blank_as_label":          # legal for ATen: a target id equal to the blank index
        T, ys = 40, [torch.tensor([4, 0, 6, 0, 0, 4]), torch.tensor([0]), torch.tensor([7, 7, 0])]
    elif case == "tb1_and_short":           # one-frame inputs, empty target, odd input lengths
        T, ys = 9, [torch.tensor([5]), torch.tensor([], dtype=torch.int64), torch.tensor([6, 2, 6]), torch.tensor([4])]
    else:                                   # runs of one label (no skip transitions anywhere)
        T, ys = 64, [torch.full((20,), 17), torch.full((31,), 200)]
    B = len(ys)
    logits = (torch.randn(T, B, C, generator=g) * 3).to(dev)
    in_lens = torch.tensor([T - 2 * b for b in range(B)], dtype=torch.int64)
    if case == "tb1_and_short":
        in_lens = torch.tensor([1, 1, 7, 5], dtype=torch.int64)
    targets = torch.cat(ys)
    tl = torch.tensor([len(y) for y in ys], dtype=torch.int64)
    ref, gref = _torch_ref(logits, targets, in_lens, tl, dev)
    loss, nll, grad = ctc_fwd_bwd(logits, targets, in_lens, tl)
    assert abs(float(loss) - float(ref)) <= 1e-5 * abs(float(ref)) + 1e-7
    scale = float(gref.abs().max())
    gtol = max(2e-4, 8 * 1.2e-7 * float(nll.abs().max()))
    assert float((grad - gref).abs().max()) <= gtol * scale + 1e-9
    for b in range(B):
        assert float(grad[int(in_lens[b]):, b].abs().max() if int(in_lens[b]) < T else 0.0) == 0.0


def test_ctc_full_batch_properties_at_baseline_shape(dev):
    """B = 2048 utterances at the BASELINE shape (T'=128, C=367, L+2=34): full-size property checks (gradient rows sum
    to zero, loss equals the mean of nll / length) and agreement with F.ctc_loss on a slice of the batch."""
    from metaasr_crossaccent_b200.ctc import ctc_fwd_bwd
    T, B, C, L = 128, 2048, 367, 34
    g = torch.Generator().manual_seed(3)
    logits = torch.randn(T, B, C, generator=g).to(dev)
    targets = torch.randint(1, C, (B * L,), generator=g)
    in_lens = torch.full((B,), T, dtype=torch.int64)
    tl = torch.full((B,), L, dtype=torch.int64)
    loss, nll, grad = ctc_fwd_bwd(logits, targets, in_lens, tl)
    assert abs(float(loss) - float((nll / L).mean())) <= 1e-5 * abs(float(loss))
    assert float(grad.sum(-1).abs().max()) <= 1e-5 * float(grad.abs().max()) * C
    sl = slice(1000, 1016)
    ref, gref = _torch_ref(logits[:, sl].contiguous(), targets.view(B, L)[sl].reshape(-1), in_lens[sl], tl[sl], dev)
    assert abs(float((nll[sl] / L).mean()) - float(ref)) <= 1e-5 * abs(float(ref))
    # the kernel's gradient carries 1/B of the FULL batch; the slice reference 1/16
    gs = grad[:, sl].double() * (B / 16.0)
    # alpha/beta are fp32 log-probabilities of magnitude ~|nll|: posteriors carry ~8 ulp(|nll|) relative error
    gtol = max(2e-4, 8 * 1.2e-7 * float(nll.abs().max()))
    assert float((gs - gref).abs().max()) <= gtol * float(gref.abs().max())
