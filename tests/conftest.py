import sys
from pathlib import Path

import pytest

REPO = Path(__file__).resolve().parent.parent
for p in (REPO, REPO / "med-ts-llm_b200"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))


# the fixtures' configs say setup.dtype = "float32", which maps to the "tf32" evaluation mode by default; the suite
# exercises the default bf16 path unless a test asks for a parity mode explicitly (monkeypatch.setenv)
import os  # noqa: E402
os.environ.setdefault("MTS_PRECISION", "bf16")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """The tests exercise the in-tree libmtsb200.so; (re)build it when sources changed and nvcc exists."""
    from medtsllm_b200 import _build
    try:
        _build.build(verbose=False)
    except RuntimeError as e:
        if "nvcc not found" in str(e) and _build.LIB_PATH.exists():
            return
        raise
