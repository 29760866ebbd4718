import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden", "pointdiff_golden.pt")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    return torch.load(GOLDEN, weights_only=True)


@pytest.fixture(scope="session")
def sd33():
    from oracle import pointdiff_oracle as O
    return O.make_synthetic_checkpoint(seed=24, alpha=1.0 / 33.0)


@pytest.fixture(scope="session")
def sd3300():
    from oracle import pointdiff_oracle as O
    return O.make_synthetic_checkpoint(seed=24, alpha=1.0 / 3300.0)


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def same_set_metrics(a: dict, b: dict) -> bool:
    """MMD-CD / COV-CD / 1-NNA-CD dicts agree: COV is a ratio of integers (exact), MMD a float32 mean (summation order differs
    between devices), 1-NNA a ratio that one side may have rounded to float32."""
    return (a["cov_cd"] == b["cov_cd"] and abs(a["mmd_cd"] - b["mmd_cd"]) <= 1e-5 * abs(b["mmd_cd"])
            and abs(a["1nna_cd"] - b["1nna_cd"]) <= 1e-6)
