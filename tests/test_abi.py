"""CPU: the C-ABI library loads, exports every symbol include/clusten_b200.h declares, the ctypes table mirrors the
header, and argument errors are reported without touching a GPU."""
import ctypes
import os
import re

import pytest
import torch

from autofocusformermod_b200 import _build, _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "clusten_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(int|size_t|long long|const char \*)\s*(clusten_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        args = [a.strip() for a in m.group(3).split(",")]
        out[m.group(2)] = 0 if args == ["void"] else len(args)
    return out


@pytest.fixture(scope="module")
def lib():
    _build.build()
    return _lib.lib()


def test_every_declared_symbol_is_exported(lib):
    decl = _declared()
    assert len(decl) >= 18
    for name in decl:
        assert hasattr(lib, name), f"{name} declared in include/clusten_b200.h but not exported"


def test_ctypes_table_mirrors_header(lib):
    decl = _declared()
    assert set(decl) == set(_lib.SIGNATURES)
    for name, nargs in decl.items():
        assert len(_lib.SIGNATURES[name][1]) == nargs, name


def test_abi_version(lib):
    assert lib.clusten_abi_version() == 6


def test_argument_errors_without_gpu(lib):
    # H = 0 -> CLUSTEN_EINVAL before any CUDA call
    rc = lib.clusten_qk_fwd(1, 1, 1, None, 1, 1, 0, 4, 4, 8, 4, 0, 0, 0, 0, 0, 0, 0, None)
    assert rc == -1 and b"bad sizes" in lib.clusten_last_error()
    rc = lib.clusten_qk_fwd(1, 1, 1, None, 1, 1, 1, 4, 4, 8, 4, 0, 0, 0, 0, 0, 0, 7, None)
    assert rc == -2                                  # unknown dtype
    rc = lib.clusten_knn(1, 1, 1, 4, 4, 17, 1, None, None)
    assert rc == -3                                  # k > 16
    rc = lib.clusten_csr_build(1, 1, 4, 300, 4, 1, 1, 1, 1 << 30, None, None)
    assert rc == -3                                  # M > 256
    rc = lib.clusten_sfc_cluster(1, 1, 16, 8, 4, 4, 1, 1, 1, None, 1, 1, 0, None)
    assert rc == -4                                  # workspace too small
    assert lib.clusten_csr_workspace_bytes(2, 64, 48, 64) >= 4 * 2 * 64 * 48 * 4


def test_argument_errors_of_the_structure_builders(lib):
    """Plans / packs / helper kernels reject bad arguments before any CUDA call (still no GPU here)."""
    assert lib.clusten_wf_plan_bytes(2, 100, 48, 400) > lib.clusten_wf_plan_bytes(1, 100, 48, 400) >= 256
    assert lib.clusten_pack_bytes(2, 100, 48, 400) > lib.clusten_pack_bytes(1, 100, 48, 400) >= 256
    rc = lib.clusten_wf_plan_build(1, 2, 100, 9, 400, 1, 1 << 30, None)
    assert rc == -3 and b"M % 8" in lib.clusten_last_error()                 # no octet structure
    rc = lib.clusten_wf_plan_build(1, 2, 100, 48, 400, 1, 16, None)
    assert rc == -4                                                         # plan buffer too small
    rc = lib.clusten_pack_build(1, None, 2, 100, 48, 400, 1, 16, None)
    assert rc == -4
    rc = lib.clusten_col_sum(1, 1, 10, 12, 12, 2, None)
    assert rc == -3                                                         # C not a multiple of 8 for a 16-bit type
    rc = lib.clusten_blank_grad(1, 1, 1, 1, 1, 1, 2, 2, 16, 12, 0, 0, 0, 0, 0, 0, 2, None)
    assert rc == -3                                                         # C % 8 != 0
    rc = lib.clusten_table_grad(1, 1, 0, 1, 10, 0, None, 4, 10, 0, 0, 0, 0, None)
    assert rc == -1                                                         # U <= 0
    rc = lib.clusten_scale_residual_fwd(16, 16, None, None, 16, 2, 10, 6, 0, 0, 0, None)
    assert rc == -3 and b"C % 4" in lib.clusten_last_error()                 # channel count not a multiple of 4
    rc = lib.clusten_scale_residual_fwd(16, 16, None, None, 16, 2, 10, 8, 2, 0, 0, None)
    assert rc == -2                                                         # (bf16 res, fp32 x) is not a supported pair
    rc = lib.clusten_scale_residual_bwd(16, None, None, None, None, None, 2, 10, 8, 0, 0, None)
    assert rc == -1                                                         # neither d_x nor d_gamma wanted
    rc = lib.clusten_table_linear_fwd(16, 16, None, 16, 100, 9, 4, None, None)
    assert rc == -3 and b"F <= 8" in lib.clusten_last_error()                # more than 8 input features
    rc = lib.clusten_table_linear_bwd(16, 16, None, None, 100, 5, 4, None, None)
    assert rc == -1                                                         # d_weight missing
    rc = lib.clusten_linear_f32(16, 16, None, 16, 100, 48, 32, 48, 32, None)
    assert rc == -3 and b"K % 32" in lib.clusten_last_error()                # reduction length not a multiple of the k chunk
    rc = lib.clusten_linear_f32(16, 16, None, 16, 100, 64, 32, 32, 32, None)
    assert rc == -1                                                         # row stride shorter than the row


def test_ops_refuse_cpu_tensors():
    """No CPU fallback: CHECK_CUDA of clustenqk_cuda.cpp:21."""
    from autofocusformermod_b200 import CLUSTENQKFunction, WEIGHTEDGATHERFunction, knn_keops
    q = torch.randn(1, 1, 4, 4)
    idx = torch.zeros(1, 4, 2, dtype=torch.int64)
    with pytest.raises(RuntimeError, match="CUDA"):
        CLUSTENQKFunction.apply(q, q, idx)
    with pytest.raises(RuntimeError, match="CUDA"):
        WEIGHTEDGATHERFunction.apply(idx, torch.randn(1, 4, 2), torch.randn(1, 4, 4))
    with pytest.raises(RuntimeError, match="CUDA"):
        knn_keops(torch.zeros(1, 4, 2), torch.zeros(1, 4, 2), 2)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.lib()
