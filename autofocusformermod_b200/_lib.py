"""ctypes binding of libclusten_b200.so (the C ABI declared in include/clusten_b200.h).

PyTorch is plumbing only: tensors provide device memory (``data_ptr()``) and the current CUDA stream handle.
There is NO CPU path and NO fallback: if the shared library is missing or a call fails, a RuntimeError is raised
(same contract as the import guard of the reference, mask2former/modeling/clusten/clusten.py:8-16).
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libclusten_b200.so")

_c = ctypes
_P, _I, _L, _Z = _c.c_void_p, _c.c_int, _c.c_int64, _c.c_size_t

# name -> (restype, argtypes); mirrors include/clusten_b200.h one to one (tests check the two stay in sync)
SIGNATURES = {
    "clusten_abi_version": (_I, []),
    "clusten_last_error": (_c.c_char_p, []),
    "clusten_kernel_launches": (_c.c_longlong, []),
    "clusten_csr_workspace_bytes": (_Z, [_I] * 4),
    "clusten_csr_build": (_I, [_P, _I, _I, _I, _I, _P, _P, _P, _Z, _P, _P]),
    "clusten_pack_bytes": (_Z, [_I] * 4),
    "clusten_pack_build": (_I, [_P, _P, _I, _I, _I, _I, _P, _Z, _P]),
    "clusten_pack_inverse": (_I, [_P, _Z, _I, _I, _I, _I, _P]),
    "clusten_qk_fwd": (_I, [_P] * 5 + [_I] * 6 + [_L] * 6 + [_I, _P]),
    "clusten_qk_bwd": (_I, [_P] * 9 + [_I] * 6 + [_L] * 12 + [_I, _P]),
    "clusten_av_fwd": (_I, [_P] * 5 + [_I] * 6 + [_L] * 9 + [_I, _P]),
    "clusten_av_bwd": (_I, [_P] * 9 + [_I] * 6 + [_L] * 12 + [_I, _P]),
    "clusten_attn_fwd": (_I, [_P] * 13 + [_I] * 6 + [_L] * 12 + [_I, _P]),
    "clusten_attn_bwd": (_I, [_P] * 18 + [_I] * 6 + [_L] * 18 + [_I, _P]),
    "clusten_attn_pos_fwd": (_I, [_P] * 15 + [_I] * 6 + [_L] * 12 + [_I, _P]),
    "clusten_attn_pos_bwd": (_I, [_P] * 21 + [_I] * 6 + [_L] * 18 + [_I, _P]),
    "clusten_col_sum": (_I, [_P, _P, _L, _I, _L, _I, _P]),
    "clusten_linear_f32": (_I, [_P] * 4 + [_L, _I, _I, _L, _L, _P]),
    "clusten_tf32_split": (_I, [_P, _P, _P, _L, _P]),
    "clusten_f16_split": (_I, [_P, _P, _P, _L, _I, _P, _P, _P]),
    "clusten_linear_tc_f32": (_I, [_P] * 7 + [_L, _I, _I, _L, _L, _L, _I, _c.c_float, _I, _I] + [_P] * 4 + [_I, _P, _P]),
    "clusten_table_linear_fwd": (_I, [_P] * 4 + [_I] * 3 + [_P, _P]),
    "clusten_table_linear_bwd": (_I, [_P] * 4 + [_I] * 3 + [_P, _P]),
    "clusten_scale_residual_fwd": (_I, [_P] * 5 + [_L, _L, _I, _I, _I, _I, _P]),
    "clusten_scale_residual_bwd": (_I, [_P] * 6 + [_L, _L, _I, _I, _I, _P]),
    "clusten_blank_grad": (_I, [_P] * 6 + [_I] * 4 + [_L] * 6 + [_I, _P]),
    "clusten_scatter_rows": (_I, [_P] * 6 + [_I] * 6 + [_L] * 9 + [_I, _P]),
    "clusten_layer_norm_fwd": (_I, [_P] * 6 + [_L, _I, _c.c_float, _I, _I, _P]),
    "clusten_layer_norm_bwd": (_I, [_P] * 8 + [_L, _I, _I, _I, _P]),
    "clusten_prepare_workspace_bytes": (_Z, []),
    "clusten_stage_prepare": (_I, [_P] * 4 + [_I] * 5 + [_P] * 6 + [_I, _P, _P, _Z, _P]),
    "clusten_table_rank": (_I, [_P, _L, _P, _P, _I, _P, _P, _Z, _P]),
    "clusten_rel_pos_features": (_I, [_P, _P, _L, _P]),
    "clusten_table_gather": (_I, [_P, _P, _I, _P, _L, _I, _I, _I, _P]),
    "clusten_table_grad": (_I, [_P, _P, _I, _P, _L, _I, _P, _I, _L, _L, _L, _L, _I, _P]),
    "clusten_wf_plan_bytes": (_Z, [_I] * 4),
    "clusten_wf_plan_build": (_I, [_P, _I, _I, _I, _I, _P, _Z, _P]),
    "clusten_wf_fwd": (_I, [_P] * 5 + [_I] * 6 + [_L] * 2 + [_I, _P]),
    "clusten_wf_bwd": (_I, [_P] * 9 + [_I] * 6 + [_L] * 4 + [_I, _P]),
    "clusten_wg_fwd": (_I, [_P] * 4 + [_I] * 5 + [_L] * 2 + [_I, _P]),
    "clusten_wg_bwd": (_I, [_P] * 8 + [_I] * 5 + [_L] * 4 + [_I, _P]),
    "clusten_msdetrpc_fwd": (_I, [_P] * 5 + [_I] * 6 + [_L] * 2 + [_I, _P]),
    "clusten_msdetrpc_bwd": (_I, [_P] * 10 + [_I] * 6 + [_L] * 4 + [_I, _P]),
    "clusten_knn": (_I, [_P, _P, _I, _I, _I, _I, _P, _P, _P]),
    "clusten_sfc_workspace_bytes": (_Z, [_I, _I]),
    "clusten_sfc_cluster": (_I, [_P, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _Z, _P]),
    "clusten_topk_workspace_bytes": (_Z, [_I, _I]),
    "clusten_topk_select": (_I, [_P, _I, _I, _I, _P, _L, _P, _Z, _P]),
    "clusten_mask_select": (_I, [_P, _I, _I, _I, _P, _L, _P]),
    "clusten_merge_scores": (_I, [_P, _P, _I, _P, _c.c_float, _I, _I, _P, _P, _I, _I, _P]),
    "clusten_gather_rows": (_I, [_P, _P, _P, _I, _I, _I, _I, _P, _P]),
    "clusten_stem_conv_bn_gelu": (_I, [_P] * 7 + [_c.c_float, _P] + [_I] * 6 + [_P]),
    "clusten_stem_im2col": (_I, [_P, _P] + [_I] * 5 + [_P]),
}

_lib = None


def lib():
    """Load (once) and return the shared library; raises RuntimeError when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"Could not load the CLUSTEN B200 CUDA library ({LIB_PATH} is missing). Build it with "
                "`python -c 'import __graft_entry__ as g; g.build()'` (nvcc, sm_100a). There is no CPU fallback.")
        try:
            L = ctypes.CDLL(LIB_PATH)
        except OSError as e:  # pragma: no cover
            raise RuntimeError(f"Could not load the CLUSTEN B200 CUDA library: {e}") from e
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


DTYPES = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2}


def dtype_code(t):
    try:
        return DTYPES[t.dtype]
    except KeyError:
        raise RuntimeError(f"CLUSTEN B200 kernels support float32/float16/bfloat16, got {t.dtype}") from None


def check(rc, what):
    if rc != 0:
        msg = lib().clusten_last_error().decode(errors="replace")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")


def stream_ptr(device):
    return torch.cuda.current_stream(device).cuda_stream


def require_cuda(*tensors):
    """CHECK_CUDA of the reference bindings (clustenqk_cuda.cpp:21): every operand must live on one CUDA device."""
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("CLUSTEN B200 ops need CUDA tensors (there is no CPU implementation in this package)")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise RuntimeError(f"tensors on different devices: {dev} vs {t.device}")
    return dev


def ptr(t):
    return 0 if t is None else t.data_ptr()
