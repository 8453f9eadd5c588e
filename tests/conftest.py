import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "cross-modality-minipig-gan_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def _no_tf32():
    # the torch references must be true fp32 (cuDNN/cuBLAS default to TF32 for convs on Ampere+)
    try:
        import torch
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
    except Exception:  # noqa: BLE001
        pass


_no_tf32()


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")
    config.addinivalue_line("markers", "slow: long CPU test (set MPGAN_SLOW=1 to run)")


def rel_l2(a, b):
    """||a-b|| / ||b|| in fp64 (b is the reference)."""
    import torch
    a, b = a.detach().double().cpu().flatten(), b.detach().double().cpu().flatten()
    den = float(torch.linalg.norm(b))
    num = float(torch.linalg.norm(a - b))
    return num / den if den > 0 else num


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
