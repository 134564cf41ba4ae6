import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


def rel_err(a, b):
    """max |a-b| / max |b| -- the comparator of SURVEY.md 8(c)."""
    import torch
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    den = b.abs().max().clamp_min(1e-30)
    return float((a - b).abs().max() / den)


@pytest.fixture(params=["per-warp", "tma-staged"])
def attn_fwd_kernel(request, monkeypatch):
    """Runs a test once with each forward kernel behind clusten_attn_fwd / clusten_attn_pos_fwd: the per-warp kernel
    (clusten_fused.cu, default) and the CTA-cooperative TMA-staged kernel (clusten_fused_tma.cu, CLUSTEN_TMA_ATTN=1)."""
    monkeypatch.setenv("CLUSTEN_TMA_ATTN", "1" if request.param == "tma-staged" else "0")
    return request.param
