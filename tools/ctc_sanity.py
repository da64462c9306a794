#!/usr/bin/env python3
"""A few small CTC launches (BASELINE frame / class counts, ragged lengths, many repeated labels, a long target, a
short input) for quick checks under a debugger or a sanitizer; prints loss, nll and the gradient mass."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from metaasr_crossaccent_b200.ctc import ctc_fwd_bwd  # noqa: E402

g = torch.Generator().manual_seed(0)
T, B, C = 128, 4, 367
lg = torch.randn(T, B, C, generator=g).cuda()
ys = [torch.randint(1, C, (34,), generator=g), torch.randint(1, 5, (20,), generator=g), torch.randint(1, C, (90,), generator=g),
      torch.randint(1, C, (7,), generator=g)]
il = torch.tensor([128, 101, 64, 5])
tl = torch.tensor([len(y) for y in ys])
loss, nll, grad = ctc_fwd_bwd(lg, torch.cat(ys), il, tl)
torch.cuda.synchronize()
print("loss", float(loss), "nll", nll.tolist(), "grad abs sum", float(grad.abs().sum()))
