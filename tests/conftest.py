import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")
    # the torch ops used as fp32 references in the -m gpu tests must be real fp32 (cuDNN convolutions
    # default to TF32, which is ~1e-3 relative)
    import torch
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


@pytest.fixture(scope="session")
def golden_dir():
    return ROOT / "tests" / "golden"
