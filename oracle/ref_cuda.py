"""Loader for the reference's OWN CUDA kernels built into oracle/_ref by oracle/build_ref.py (TEST INFRASTRUCTURE).

The four pybind modules are compiled, unmodified, from the sources under /root/reference for sm_100a in the authoring
container; only the built ``.so`` files travel to the GPU box.  ``functions()`` wraps them exactly as the reference's
``clusten.py:19-120`` does (contiguous copies, K transposed for the QK forward, dtype casts) so that parity tests and
the "reference kernel" timing call the real thing.
"""
import importlib.util
import os

import torch

_REF = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
_mods = {}


def available():
    return all(os.path.exists(os.path.join(_REF, m, m + ".so"))
               for m in ("clustenqk_cuda", "clustenav_cuda", "clustenwf_cuda", "weighted_gather_cuda"))


def module(name):
    if name not in _mods:
        path = os.path.join(_REF, name, name + ".so")
        spec = importlib.util.spec_from_file_location(name, path)
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        _mods[name] = m
    return _mods[name]


def qk_forward(query, key, nbhd_idx):                                       # clusten.py:24-35
    query, key = query.contiguous(), key.contiguous().to(query.dtype)
    return module("clustenqk_cuda").forward(query, key.permute(0, 1, 3, 2).contiguous(), nbhd_idx.contiguous())


def qk_backward(grad_attn, query, key, nbhd_idx):                           # clusten.py:38-42
    return module("clustenqk_cuda").backward(grad_attn.contiguous(), query.contiguous(), key.contiguous(), nbhd_idx.contiguous())


def av_forward(attn, v, nbhd_idx):                                          # clusten.py:50-61
    return module("clustenav_cuda").forward(attn.contiguous(), v.contiguous().to(attn.dtype), nbhd_idx.contiguous())


def av_backward(grad_feat, attn, v, nbhd_idx):                              # clusten.py:64-68
    return module("clustenav_cuda").backward(grad_feat.contiguous(), attn.contiguous(), v.contiguous(), nbhd_idx.contiguous())


def wf_forward(weights, feat, nbhd_idx):                                    # clusten.py:76-87
    return module("clustenwf_cuda").forward(weights.contiguous(), feat.contiguous().to(weights.dtype), nbhd_idx.contiguous())


def wf_backward(grad_out, weights, feat, nbhd_idx):                         # clusten.py:90-94
    return module("clustenwf_cuda").backward(grad_out.contiguous(), weights.contiguous(), feat.contiguous(), nbhd_idx.contiguous())


def wg_forward(nbhd_idx, weights, feat):                                    # clusten.py:102-113
    return module("weighted_gather_cuda").forward(nbhd_idx.contiguous(), weights.contiguous().to(feat.dtype), feat.contiguous())


def wg_backward(grad_out, nbhd_idx, weights, feat):                         # clusten.py:116-120
    return module("weighted_gather_cuda").backward(grad_out.contiguous(), nbhd_idx.contiguous(), weights.contiguous(), feat.contiguous())
