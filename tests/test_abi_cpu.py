"""CPU: the C-ABI shared library loads and exports every symbol include/metaasr_b200.h declares,
and the product path fails loudly (no fallback) when there is no GPU."""
import re
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]


def declared_symbols():
    text = (ROOT / "include" / "metaasr_b200.h").read_text()
    return sorted(set(re.findall(r"\b(masr_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from metaasr_crossaccent_b200 import _lib
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    bound = set(_lib.SIGNATURES) | set(_lib.EXTRA)
    assert set(names) == bound, (set(names) ^ bound)
    assert lib.masr_abi_version() == 3


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only check")
def test_no_cpu_fallback():
    from metaasr_crossaccent_b200 import _lib
    from metaasr_crossaccent_b200.ops import CudaBackend
    with pytest.raises(_lib.MetaASRLibraryError):
        CudaBackend("cpu")
    lib = _lib.load()
    assert lib.masr_init(0) == -3                      # MASR_E_NOGPU
    assert b"no CPU fallback" in lib.masr_last_error()
