"""CPU: the arithmetic of csrc/posbias.cuh (relative-position bias computed from positions, used by the opt-in
clusten_attn_pos_* kernels) compiled FOR THE HOST and pinned against the reference's table formulation
(aff.py:17-31 pre_table, :129-132 pos_embed(pre_table)[pe_idx], :481-485 pe_idx from positions) as restated by the oracle."""
import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest
import torch

from oracle import inputs
from oracle import point_ops as pt

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def host_lib(tmp_path_factory):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    out = str(tmp_path_factory.mktemp("posbias") / "libposbias_host.so")
    src = os.path.join(ROOT, "tests", "native", "posbias_host.cu")
    r = subprocess.run([nvcc, "-O2", "-std=c++17", "-shared", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-o", out, src],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    L = ctypes.CDLL(out)
    fp = ctypes.POINTER(ctypes.c_float)
    L.posbias_host_bias.argtypes = [fp, fp, fp, fp, ctypes.c_int, ctypes.c_int, fp]
    L.posbias_host_grad.argtypes = [fp, fp, fp, ctypes.c_int, fp]
    return L


def _p(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


@pytest.mark.parametrize("n,hw", [(1024, 64), (655, 128), (300, 512)])
def test_position_bias_equals_table_formulation(host_lib, n, hw):
    B, H = 2, 3
    pos, idx, _, pe_idx = inputs.structured_neighbourhood(B, n, hw, hw, 8, 48, seed=n)
    M = idx.shape[-1]
    g = torch.Generator().manual_seed(n)
    W, b = torch.randn(H, 5, generator=g) * 0.2, torch.randn(H, generator=g)
    table = pt.build_pre_table()                                                       # aff.py:17-31, [1023^2, 5]
    want = (table[pe_idx.reshape(-1)] @ W.t() + b).reshape(B, n, M, H)                 # pos_embed(pre_table)[pe_idx]
    q = pos.float().unsqueeze(2).expand(-1, -1, M, -1).reshape(-1, 2).contiguous().numpy()
    k = pos.float().gather(1, idx.reshape(B, -1, 1).expand(-1, -1, 2)).reshape(-1, 2).contiguous().numpy()
    got = np.empty((q.shape[0], H), dtype=np.float32)
    Wn, bn = W.contiguous().numpy(), b.contiguous().numpy()
    host_lib.posbias_host_bias(_p(q), _p(k), _p(Wn), _p(bn), q.shape[0], H, _p(got))
    got = torch.from_numpy(got).reshape(B, n, M, H)
    assert float((got - want).abs().max()) <= 2e-6 * float(want.abs().max())
    # gradient of pos_embed: sum ds * [feat | 1] against autograd through the table formulation
    ds = torch.randn(B * n * M, generator=g)
    Wg, bg = W.clone().requires_grad_(True), b.clone().requires_grad_(True)
    ((table[pe_idx.reshape(-1)] @ Wg.t() + bg)[:, 0] * ds).sum().backward()
    grad = np.empty(6, dtype=np.float32)
    dsn = ds.contiguous().numpy()
    host_lib.posbias_host_grad(_p(q), _p(k), _p(dsn), q.shape[0], _p(grad))
    ref = torch.cat([Wg.grad[0], bg.grad[:1]])
    assert float((torch.from_numpy(grad) - ref).abs().max()) <= 1e-3 * float(ref.abs().max())      # fp32 running sum of ~50 k terms


def test_position_bias_clamps_like_the_table_index(host_lib):
    """Offsets beyond +-511 clamp to the table border (aff.py:484) and the centre row is the zeroed one (aff.py:31)."""
    q = np.array([[0, 0], [0, 0], [600, 10], [5, 5]], dtype=np.float32)
    k = np.array([[700, -3], [0, 0], [0, 900], [5, 5]], dtype=np.float32)
    W = np.array([[1, 10, 100, 1000, 10000]], dtype=np.float32)
    b = np.array([0.5], dtype=np.float32)
    got = np.empty((4, 1), dtype=np.float32)
    host_lib.posbias_host_bias(_p(q), _p(k), _p(W), _p(b), 4, 1, _p(got))
    table = pt.build_pre_table()
    rel = torch.from_numpy(k) - (torch.from_numpy(q) - 511)
    rel = rel.clamp(0, 1022).long()
    want = table[rel[:, 1] * 1023 + rel[:, 0]] @ torch.from_numpy(W).t() + 0.5
    assert float((torch.from_numpy(got) - want).abs().max()) <= 1e-5 * float(want.abs().max())
    assert got[1, 0] == 0.5 and got[3, 0] == 0.5                                       # centre: all five features are zero
