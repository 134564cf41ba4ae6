"""The CLUSTEN autograd boundary: drop-in replacements of the reference's ``torch.autograd.Function``s
(mask2former/modeling/clusten/clusten.py:19-120) on top of libclusten_b200.so.

Same class names, same ``.apply`` argument order, shapes, dtype-cast rules and ``None``-gradient positions:

    CLUSTENQKFunction.apply(query[B,H,N,C], key[B,H,N,C], nbhd_idx[B,N,M])        -> attn[B,H,N,M]      (clusten.py:24)
    CLUSTENAVFunction.apply(attn[B,H,N,M], v[B,H,N,C], nbhd_idx[B,N,M])           -> feat[B,H,N,C]      (clusten.py:50)
    CLUSTENWFFunction.apply(weights[B,N',M,IC], feat[B,N,C], nbhd_idx[B,N',M])    -> feat_new[B,N',IC,C](clusten.py:76)
    WEIGHTEDGATHERFunction.apply(nbhd_idx[B,N,K], weights[B,N,K], feat[B,N_,C])   -> feat_new[B,N,C]    (clusten.py:102)

Differences from the reference that a caller can observe, all deliberate:
  * row operands (query/key/v/feat) may be arbitrary strided views with unit innermost stride -- the permuted
    ``b h n c`` views of aff.py:111-113 are consumed in place instead of being copied three times (clusten.py:25-33);
  * bfloat16 is supported (the reference dispatches fp64/fp32/fp16 only); accumulation is always fp32; float64 is not;
  * backward is deterministic: the ``fastAtomicAdd`` scatter is replaced by a gather over an inverse neighbour list
    that is built once per index tensor and cached ON that tensor (``nbhd_idx`` is the same object for every block of
    an AFF stage, aff.py:487-493);
  * ``feat`` of AV is returned as a [B,H,N,C] VIEW of token-major [B,N,H,C] memory so that the caller's
    ``permute(0,2,1,3).reshape(b,n,c)`` (aff.py:154) is free.  Set ``TOKEN_MAJOR_AV_OUTPUT = False`` for a
    contiguous [B,H,N,C] result.
There is no CPU path: non-CUDA operands raise RuntimeError, like CHECK_CUDA in clustenqk_cuda.cpp:21.
"""
import os

import torch
from torch.autograd import Function

from . import _lib

TOKEN_MAJOR_AV_OUTPUT = True

# ---- optional per-kernel device timing (bench.py uses it for the roofline line) --------------------------------------
_timer = None     # dict(name=<entry point>, events=[(start, end), ...]) or None


def start_kernel_timer(name):
    """Record a (start, end) CUDA-event pair around every call of C-ABI entry point ``name`` ("*" = all of them) on
    its launch stream."""
    global _timer
    _timer = {"name": name, "events": []}


def stop_kernel_timer():
    """Returns [(entry point, milliseconds, algorithmic bytes)] per recorded call (synchronises)."""
    global _timer
    t, _timer = _timer, None
    if t is None:
        return []
    torch.cuda.synchronize()
    return [(n, a.elapsed_time(b), nb) for n, a, b, nb in t["events"]]


_DEBUG_CAPTURE = __import__("os").environ.get("CLUSTEN_DEBUG_CAPTURE", "0") == "1"
_cudart = None


def _capture_status(tag):
    """Debug aid (CLUSTEN_DEBUG_CAPTURE=1): report the first point at which the current stream's capture turns invalid."""
    global _cudart
    if not _DEBUG_CAPTURE:
        return
    import ctypes
    if _cudart is None:
        _cudart = ctypes.CDLL("libcudart.so.12")
    st = ctypes.c_int(0)
    rc = _cudart.cudaStreamIsCapturing(ctypes.c_void_p(torch.cuda.current_stream().cuda_stream), ctypes.byref(st))
    if st.value == 2 or rc != 0:
        print(f"[capture] INVALID at {tag} (rc={rc}, status={st.value})", flush=True)
        raise RuntimeError(f"capture invalidated before/at {tag}")


_launch_count = 0


def launch_count():
    """Number of C-ABI compute entry points called so far in this process (each launches >= 1 kernel of ours)."""
    return _launch_count


def kernel_launches():
    """Number of CUDA kernels libclusten_b200.so has launched in this process (clusten_kernel_launches)."""
    return int(_lib.lib().clusten_kernel_launches())


_flops = {}                                            # entry point -> [executed, algorithmic] FLOPs so far (GEMM-shaped entry points only)


def flops_issued(name):
    """(executed, algorithmic) FLOPs issued through entry point ``name`` in this process.  Algorithmic = 2 R K N per Linear call;
    executed = the tensor-core work behind it expressed at the bf16 / fp16 dense rate: three MMAs per product with the fp16 split
    (3 x 2RKN), three TF32 MMAs at half that rate with the TF32 split (6 x 2RKN)."""
    e = _flops.get(name, (0, 0))
    return float(e[0]), float(e[1])


def _call(name, dev, *args, nbytes=0, flops=None):
    """Invoke C-ABI entry point ``name`` on the current stream of ``dev``.  ``nbytes`` = ALGORITHMIC bytes of the call
    (each operand read once, each result written once, idx as the int64 delivered) for roofline reports."""
    global _launch_count
    _launch_count += 1
    if flops:                                          # (executed in bf16-rate equivalents, algorithmic)
        e = _flops.setdefault(name, [0, 0])
        e[0] += flops[0]
        e[1] += flops[1]
    fn = getattr(_lib.lib(), name)
    _capture_status("before " + name)
    timed = _timer is not None and _timer["name"] in (name, "*")
    if timed:
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(torch.cuda.current_stream(dev))
    rc = fn(*args, _lib.stream_ptr(dev))
    if timed:
        b.record(torch.cuda.current_stream(dev))
        _timer["events"].append((name, a, b, nbytes))
    _lib.check(rc, name)


# ---- helpers ---------------------------------------------------------------------------------------------------------
def _rows(t):
    """Tensor addressed as base + sum(idx*stride) with unit innermost stride (copied only if it is not)."""
    if t.shape[-1] > 1 and t.stride(-1) != 1:
        t = t.contiguous()
    elif t.shape[-1] == 1 and t.stride(-1) != 1:
        t = t.contiguous()
    return t


def _idx(nbhd_idx):
    if nbhd_idx.dtype != torch.int64:
        raise RuntimeError(f"nbhd_idx must be int64 (got {nbhd_idx.dtype})")
    return nbhd_idx if nbhd_idx.is_contiguous() else nbhd_idx.contiguous()


def _check_shapes(cond, msg):
    if not cond:
        raise RuntimeError(msg)


def inverse_neighbour_list(nbhd_idx, Nk, with_pack=False, wf_plan_buf=None, pack_buf=None):
    """(offsets int32 [B,Nk+1], entries uint32-as-int32 [B,Nq*M]) of clusten_csr_build, cached on the index tensor.
    ``with_pack`` (QK / AV backward only): the list is built beside the tile pack and SKIPPED on the device when the
    pack routes the call to the tile-union kernels, which do not need it.  ``wf_plan_buf`` (WF backward): same, keyed
    on the WF plan's device-side flag."""
    cache = getattr(nbhd_idx, "_clusten_csr", None)
    ver = nbhd_idx._version
    pack = pack_buf if pack_buf is not None else neighbourhood_pack(nbhd_idx, Nk) if with_pack else wf_plan_buf
    kind = ("mpack" if pack_buf is not None else "pack" if with_pack and pack is not None else "wf" if pack is not None else None)
    # (a list built beside a pack / plan may have been skipped on the device: only reuse it for the same kind of call;
    # an unconditional list serves everyone)
    # (a list built beside a mask-aware pack is only valid for that very pack buffer, which the cache keeps alive)
    if (cache is not None and cache[0] == ver and cache[1] == Nk and cache[2] == nbhd_idx.data_ptr() and cache[5] in (None, kind)
            and (cache[5] != "mpack" or cache[6] is pack_buf)):
        return cache[3], cache[4]
    B, Nq, M = nbhd_idx.shape
    dev = nbhd_idx.device
    L = _lib.lib()
    offsets = torch.zeros((B, Nk + 1), dtype=torch.int32, device=dev)     # zeros: stays a valid (empty) list when the
    entries = torch.empty((B, max(Nq * M, 1)), dtype=torch.int32, device=dev)   # build is skipped on the device
    ws_bytes = L.clusten_csr_workspace_bytes(B, Nq, M, Nk)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _call("clusten_csr_build", dev, nbhd_idx.data_ptr(), B, Nq, M, Nk, offsets.data_ptr(), entries.data_ptr(),
              ws.data_ptr(), ws_bytes, _lib.ptr(pack))
    try:
        nbhd_idx._clusten_csr = (ver, Nk, nbhd_idx.data_ptr(), offsets, entries, kind, pack_buf)
    except Exception:  # pragma: no cover  (tensor subclass without __dict__)
        pass
    return offsets, entries


USE_TILE_KERNELS = True      # False: always the generic row-gather kernels (pack == NULL)


def neighbourhood_pack(nbhd_idx, Nk, inverse=False, mask=None):
    """Opaque tile pack of clusten_pack_build for this index tensor (uint8 device buffer), cached on the tensor; None
    when the tile-union kernels are switched off.  The pack decides ON THE DEVICE whether the tensor-core kernels or the
    generic ones run (no host sync); see ``pack_flags``.  ``mask`` (uint8 [B,Nq,M]; fused attention only): a mask-aware
    pack, cached separately -- masked neighbour slots are wildcards, so padded clusters stay on the tensor-core path."""
    if not USE_TILE_KERNELS:
        return None
    attr = "_clusten_pack" if mask is None else "_clusten_pack_masked"
    cache = getattr(nbhd_idx, attr, None)
    ver = nbhd_idx._version
    B, Nq, M = nbhd_idx.shape
    dev = nbhd_idx.device
    # the mask tensor ITSELF is part of the key (and is kept alive by the cache): an address + version pair can be recycled
    mver = None if mask is None else mask._version
    if (cache is not None and cache[0] == ver and cache[1] == Nk and cache[2] == nbhd_idx.data_ptr()
            and cache[5] is mask and cache[6] == mver):
        pack, has_inv = cache[3], cache[4]
    else:
        nbytes = _lib.lib().clusten_pack_bytes(B, Nq, M, Nk)
        pack, has_inv = torch.empty(nbytes, dtype=torch.uint8, device=dev), False
        with torch.cuda.device(dev):
            _call("clusten_pack_build", dev, nbhd_idx.data_ptr(), _lib.ptr(mask), B, Nq, M, Nk, pack.data_ptr(), nbytes)
    if inverse and not has_inv:                  # the inverse lists are only needed by backward: built on first use
        with torch.cuda.device(dev):
            _call("clusten_pack_inverse", dev, pack.data_ptr(), pack.numel(), B, Nq, M, Nk)
        has_inv = True
    try:
        setattr(nbhd_idx, attr, (ver, Nk, nbhd_idx.data_ptr(), pack, has_inv, mask, mver))
    except Exception:  # pragma: no cover
        pass
    return pack


USE_WF_PLAN = True          # False: WF runs without a plan (natural token order, generic d_f)


def wf_plan(nbhd_idx, Nk):
    """Opaque WF plan of clusten_wf_plan_build for this index tensor (uint8 device buffer), cached on the tensor; None
    when M is not a multiple of 8 (no octet structure to exploit: PointConv's kNN-9, the 4-neighbour upsampling)."""
    B, Nq, M = nbhd_idx.shape
    if not USE_WF_PLAN or M % 8 != 0 or B * Nq == 0:
        return None
    cache = getattr(nbhd_idx, "_clusten_wf_plan", None)
    ver = nbhd_idx._version
    if cache is not None and cache[0] == ver and cache[1] == Nk and cache[2] == nbhd_idx.data_ptr():
        return cache[3]
    dev = nbhd_idx.device
    nbytes = _lib.lib().clusten_wf_plan_bytes(B, Nq, M, Nk)
    plan = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _call("clusten_wf_plan_build", dev, nbhd_idx.data_ptr(), B, Nq, M, Nk, plan.data_ptr(), nbytes)
    try:
        nbhd_idx._clusten_wf_plan = (ver, Nk, nbhd_idx.data_ptr(), plan)
    except Exception:  # pragma: no cover
        pass
    return plan


def wf_plan_flags(nbhd_idx, Nk):
    """(generic_d_f, impure_slots, longest_octet_list) of the WF plan -- synchronises; for tests / diagnostics."""
    plan = wf_plan(nbhd_idx, Nk)
    if plan is None:
        return (1, 0, 0)
    return tuple(int(x) for x in plan[:12].view(torch.int32).tolist())


def pack_flags(nbhd_idx, Nk, mask=None):
    """(generic_path, max_union, impure_slots, tiles_over_limit) of the pack -- synchronises; for tests / diagnostics."""
    pack = neighbourhood_pack(nbhd_idx, Nk, mask=mask)
    if pack is None:
        return (1, 0, 0, 0)
    return tuple(int(x) for x in pack[:28].view(torch.int32).tolist())


def _s3(t):
    return t.stride(0), t.stride(1), t.stride(2)


# ---- QK --------------------------------------------------------------------------------------------------------------
class CLUSTENQKFunction(Function):
    """query times key: attn[b,h,i,j] = sum_c query[b,h,i,c] * key[b,h,nbhd_idx[b,i,j],c]   (clusten.py:19-42)"""

    @staticmethod
    def forward(ctx, query, key, nbhd_idx):
        dev = _lib.require_cuda(query, key, nbhd_idx)
        if key.dtype != query.dtype:
            key = key.to(query.dtype)                                   # clusten.py:27-28
        _check_shapes(query.dim() == 4 and key.dim() == 4 and nbhd_idx.dim() == 3, "QK: query/key 4-D, nbhd_idx 3-D")
        B, H, Nq, C = query.shape
        Nk = key.shape[2]
        M = nbhd_idx.shape[2]
        _check_shapes(key.shape[0] == B and key.shape[1] == H and key.shape[3] == C and
                      nbhd_idx.shape[0] == B and nbhd_idx.shape[1] == Nq, "QK: shape mismatch")
        query, key, nbhd_idx = _rows(query), _rows(key), _idx(nbhd_idx)
        attn = torch.empty((B, H, Nq, M), dtype=query.dtype, device=dev)
        es = query.element_size()
        if attn.numel():
            with torch.cuda.device(dev):
                _call("clusten_qk_fwd", dev, query.data_ptr(), key.data_ptr(), nbhd_idx.data_ptr(),
                      _lib.ptr(neighbourhood_pack(nbhd_idx, Nk)), attn.data_ptr(),
                      B, H, Nq, Nk, C, M, *_s3(query), *_s3(key), _lib.dtype_code(query),
                      nbytes=es * (B * H * (Nq + Nk) * C + B * H * Nq * M) + 8 * B * Nq * M)
        ctx.save_for_backward(query, key, nbhd_idx)
        return attn

    @staticmethod
    def backward(ctx, grad_attn):
        query, key, nbhd_idx = ctx.saved_tensors
        dev = query.device
        B, H, Nq, C = query.shape
        Nk, M = key.shape[2], nbhd_idx.shape[2]
        grad_attn = grad_attn.contiguous()
        if grad_attn.dtype != query.dtype:
            grad_attn = grad_attn.to(query.dtype)
        d_query = torch.empty_like(query)
        d_key = torch.empty_like(key)
        d_query, d_key = _rows(d_query), _rows(d_key)
        if Nq * M == 0 or B == 0:
            return d_query.zero_(), d_key.zero_(), None
        off, ent = inverse_neighbour_list(nbhd_idx, Nk, with_pack=True)
        with torch.cuda.device(dev):
            _call("clusten_qk_bwd", dev, grad_attn.data_ptr(), query.data_ptr(), key.data_ptr(), nbhd_idx.data_ptr(),
                  off.data_ptr(), ent.data_ptr(), _lib.ptr(neighbourhood_pack(nbhd_idx, Nk, inverse=True)), d_query.data_ptr(),
                  d_key.data_ptr(), B, H, Nq, Nk, C, M,
                  *_s3(query), *_s3(key), *_s3(d_query), *_s3(d_key), _lib.dtype_code(query),
                  nbytes=query.element_size() * (B * H * Nq * M + 2 * B * H * (Nq + Nk) * C) + 8 * B * Nq * M)
        return d_query, d_key, None


# ---- AV --------------------------------------------------------------------------------------------------------------
class CLUSTENAVFunction(Function):
    """attention times value: feat[b,h,i,c] = sum_j attn[b,h,i,j] * v[b,h,nbhd_idx[b,i,j],c]   (clusten.py:45-68)"""

    @staticmethod
    def forward(ctx, attn, v, nbhd_idx):
        dev = _lib.require_cuda(attn, v, nbhd_idx)
        if attn.dtype != v.dtype:
            v = v.to(attn.dtype)                                        # clusten.py:54-55
        _check_shapes(attn.dim() == 4 and v.dim() == 4 and nbhd_idx.dim() == 3, "AV: attn/v 4-D, nbhd_idx 3-D")
        B, H, Nq, M = attn.shape
        Nk, C = v.shape[2], v.shape[3]
        _check_shapes(v.shape[0] == B and v.shape[1] == H and tuple(nbhd_idx.shape) == (B, Nq, M), "AV: shape mismatch")
        attn, v, nbhd_idx = _rows(attn), _rows(v), _idx(nbhd_idx)
        if TOKEN_MAJOR_AV_OUTPUT:
            feat = torch.empty((B, Nq, H, C), dtype=attn.dtype, device=dev).permute(0, 2, 1, 3)
        else:
            feat = torch.empty((B, H, Nq, C), dtype=attn.dtype, device=dev)
        if feat.numel():
            with torch.cuda.device(dev):
                _call("clusten_av_fwd", dev, attn.data_ptr(), v.data_ptr(), nbhd_idx.data_ptr(),
                      _lib.ptr(neighbourhood_pack(nbhd_idx, Nk)), feat.data_ptr(),
                      B, H, Nq, Nk, C, M, *_s3(attn), *_s3(v), *_s3(feat), _lib.dtype_code(attn),
                      nbytes=attn.element_size() * (B * H * Nq * M + B * H * (Nq + Nk) * C) + 8 * B * Nq * M)
        ctx.save_for_backward(attn, v, nbhd_idx)
        return feat

    @staticmethod
    def backward(ctx, grad_feat):
        attn, v, nbhd_idx = ctx.saved_tensors
        dev = attn.device
        B, H, Nq, M = attn.shape
        Nk, C = v.shape[2], v.shape[3]
        grad_feat = _rows(grad_feat)
        if grad_feat.dtype != attn.dtype:
            grad_feat = grad_feat.to(attn.dtype)
        d_attn = torch.empty((B, H, Nq, M), dtype=attn.dtype, device=dev)
        d_v = _rows(torch.empty_like(v))
        if Nq * M == 0 or B == 0:
            return d_attn, d_v.zero_(), None
        off, ent = inverse_neighbour_list(nbhd_idx, Nk, with_pack=True)
        with torch.cuda.device(dev):
            _call("clusten_av_bwd", dev, grad_feat.data_ptr(), attn.data_ptr(), v.data_ptr(), nbhd_idx.data_ptr(),
                  off.data_ptr(), ent.data_ptr(), _lib.ptr(neighbourhood_pack(nbhd_idx, Nk, inverse=True)), d_attn.data_ptr(),
                  d_v.data_ptr(), B, H, Nq, Nk, C, M,
                  *_s3(grad_feat), *_s3(attn), *_s3(v), *_s3(d_v), _lib.dtype_code(attn),
                  nbytes=attn.element_size() * (2 * B * H * Nq * M + B * H * (Nq + 2 * Nk) * C) + 8 * B * Nq * M)
        return d_attn, d_v, None


# ---- fused attention core (module-level fast path) -------------------------------------------------
def cluster_attention_fused(q, key, v, nbhd_idx, bias_tab, bias_idx, mask, blank_k, blank_v, need_probs=False):
    """Forward of the ClusterAttention core (aff.py:114-155) in one kernel: softmax over the M neighbour logits
    (q.k + bias_tab[bias_idx, h] + mask) and the blank-token logit, times v (+ blank_v).  No autograd: inference path.

    q, key, v: [B,H,N,C] (any strides, unit inner stride; q already scaled); nbhd_idx int64 [B,N,M]; bias_tab fp32
    [R,H]; bias_idx int32 [B,N,M]; mask uint8 [B,N,M] or None; blank_k / blank_v [H*C].  Returns out as a [B,N,H*C]
    tensor (token-major, what ``proj`` consumes) and, when ``need_probs``, the fp32 softmax output [B,H,N,M+1]."""
    dev = _lib.require_cuda(q, key, v, nbhd_idx, bias_tab, bias_idx, mask, blank_k, blank_v)
    B, H, Nq, C = q.shape
    Nk, M = key.shape[2], nbhd_idx.shape[2]
    dt = q.dtype
    q, key, v, nbhd_idx = _rows(q), _rows(key.to(dt)), _rows(v.to(dt)), _idx(nbhd_idx)
    bias_tab = bias_tab.to(torch.float32).contiguous()
    _check_shapes(bias_tab.dim() == 2 and bias_tab.shape[1] == H and tuple(bias_idx.shape) == (B, Nq, M) and
                  bias_idx.dtype == torch.int32 and bias_idx.is_contiguous(), "fused attention: bias table / index mismatch")
    if mask is not None:
        _check_shapes(mask.dtype == torch.uint8 and tuple(mask.shape) == (B, Nq, M) and mask.is_contiguous(), "fused attention: mask must be uint8 [B,N,M]")
    blank_k, blank_v = blank_k.to(dt).contiguous(), blank_v.to(dt).contiguous()
    out = torch.empty((B, Nq, H, C), dtype=dt, device=dev)
    ov = out.permute(0, 2, 1, 3)
    probs = torch.empty((B, H, Nq, M + 1), dtype=torch.float32, device=dev) if need_probs else None
    if out.numel():
        with torch.cuda.device(dev):
            _call("clusten_attn_fwd", dev, q.data_ptr(), key.data_ptr(), v.data_ptr(), nbhd_idx.data_ptr(),
                  _lib.ptr(neighbourhood_pack(nbhd_idx, Nk, mask=mask)), bias_tab.data_ptr(), bias_idx.data_ptr(), _lib.ptr(mask),
                  blank_k.data_ptr(), blank_v.data_ptr(), out.data_ptr(), _lib.ptr(probs), 0, B, H, Nq, Nk, C, M,
                  *_s3(q), *_s3(key), *_s3(v), *_s3(ov), _lib.dtype_code(q),
                  nbytes=q.element_size() * (B * H * (2 * Nq + 2 * Nk) * C) + 4 * B * Nq * M + 8 * B * Nq * M)
    out = out.reshape(B, Nq, H * C)
    return (out, probs) if need_probs else out


class ClusterAttentionCoreFunction(Function):
    """The ClusterAttention core (aff.py:114-155) as ONE differentiable op for fp16 / bf16 training:

        out = softmax([q.k_nbhd + bias_tab[bias_idx] + mask, q.blank_k]) @ [v_nbhd; blank_v]

    q [B,N,H,C] and kv [B,N,H,2,C] are the token-major outputs of the ``q`` / ``kv`` Linear layers (q already scaled),
    bias_tab [R,H] = ``pos_embed`` evaluated on the R referenced table rows, bias_idx int32 [B,N,M] the matching inverse
    map, mask uint8 [B,N,M] or None, blank_k / blank_v [H*C].  Returns out [B,N,H*C].  Forward = clusten_attn_fwd (saves
    only out and the log-sum-exp); backward = clusten_attn_bwd + two clusten_scatter_rows + clusten_table_grad: the
    [B,H,N,M+1] fp32 tensors autograd keeps for the reference's glue passes never exist."""

    @staticmethod
    def forward(ctx, q, kv, bias_tab, blank_k, blank_v, nbhd_idx, bias_idx, mask, count=None):
        dev = _lib.require_cuda(q, kv, bias_tab, blank_k, blank_v, nbhd_idx, bias_idx, mask)
        ctx.count = count                  # device int32 scalar: bias-table rows actually referenced (see table_lookup)
        _check_shapes(q.dim() == 4 and kv.dim() == 5 and kv.shape[3] == 2 and q.dtype == kv.dtype and
                      q.dtype in (torch.float16, torch.bfloat16), "fused attention: q [B,N,H,C], kv [B,N,H,2,C], fp16/bf16")
        B, N, H, C = q.shape
        M = nbhd_idx.shape[2]
        dt = q.dtype
        q, kv, nbhd_idx = q.contiguous(), kv.contiguous(), _idx(nbhd_idx)
        tab = bias_tab.detach().to(torch.float32).contiguous()
        bk, bv = blank_k.detach().to(dt).contiguous(), blank_v.detach().to(dt).contiguous()
        _check_shapes(tab.dim() == 2 and tab.shape[1] == H and tuple(bias_idx.shape) == (B, N, M) and
                      bias_idx.dtype == torch.int32 and bias_idx.is_contiguous(), "fused attention: bias table / index mismatch")
        if mask is not None:
            _check_shapes(mask.dtype == torch.uint8 and tuple(mask.shape) == (B, N, M) and mask.is_contiguous(), "fused attention: mask must be uint8 [B,N,M]")
        out = torch.empty((B, N, H, C), dtype=dt, device=dev)
        lse = torch.empty((B, H, N), dtype=torch.float32, device=dev)
        qv, kk, vv, ov = q.permute(0, 2, 1, 3), kv[:, :, :, 0].permute(0, 2, 1, 3), kv[:, :, :, 1].permute(0, 2, 1, 3), out.permute(0, 2, 1, 3)
        if out.numel():
            with torch.cuda.device(dev):
                _call("clusten_attn_fwd", dev, qv.data_ptr(), kk.data_ptr(), vv.data_ptr(), nbhd_idx.data_ptr(),
                      _lib.ptr(neighbourhood_pack(nbhd_idx, N, mask=mask)), tab.data_ptr(), bias_idx.data_ptr(), _lib.ptr(mask),
                      bk.data_ptr(), bv.data_ptr(), out.data_ptr(), 0, lse.data_ptr(), B, H, N, N, C, M,
                      *_s3(qv), *_s3(kk), *_s3(vv), *_s3(ov), _lib.dtype_code(q),
                      nbytes=q.element_size() * (B * H * 4 * N * C) + 4 * B * N * M + 8 * B * N * M)
        ctx.save_for_backward(q, kv, tab, bk, bv, out, lse, nbhd_idx, bias_idx, mask)
        ctx.meta = (bias_tab.dtype, blank_k.dtype, blank_v.dtype)
        return out.view(B, N, H * C)

    @staticmethod
    def backward(ctx, d_out):
        q, kv, tab, bk, bv, out, lse, nbhd_idx, bias_idx, mask = ctx.saved_tensors
        dev = q.device
        B, N, H, C = q.shape
        M = nbhd_idx.shape[2]
        dt = q.dtype
        d_out = d_out.to(dt).contiguous().view(B, N, H, C)
        d_q = torch.empty_like(q)
        d_kv = torch.empty_like(kv)
        P = torch.empty((B, H, N, M), dtype=dt, device=dev)
        dS = torch.empty((B, H, N, M), dtype=dt, device=dev)
        Pb = torch.empty((B, H, N), dtype=torch.float32, device=dev)
        dSb = torch.empty((B, H, N), dtype=torch.float32, device=dev)
        d_tab = torch.zeros(tab.shape, dtype=torch.float32, device=dev)
        if q.numel():
            hv = lambda t: t.permute(0, 2, 1, 3)
            qv, kk, vv, ov, gv, dqv = hv(q), hv(kv[:, :, :, 0]), hv(kv[:, :, :, 1]), hv(out), hv(d_out), hv(d_q)
            dkv, dvv = hv(d_kv[:, :, :, 0]), hv(d_kv[:, :, :, 1])
            pack = neighbourhood_pack(nbhd_idx, N, inverse=True, mask=mask)
            off, ent = inverse_neighbour_list(nbhd_idx, N, with_pack=True, pack_buf=pack if mask is not None else None)
            code = _lib.dtype_code(q)
            es = q.element_size()
            with torch.cuda.device(dev):
                _call("clusten_attn_bwd", dev, gv.data_ptr(), ov.data_ptr(), lse.data_ptr(), qv.data_ptr(), kk.data_ptr(), vv.data_ptr(),
                      nbhd_idx.data_ptr(), _lib.ptr(pack), tab.data_ptr(), bias_idx.data_ptr(), _lib.ptr(mask), bk.data_ptr(),
                      bv.data_ptr(), d_q.data_ptr(), P.data_ptr(), dS.data_ptr(), Pb.data_ptr(), dSb.data_ptr(),
                      B, H, N, N, C, M, *_s3(qv), *_s3(kk), *_s3(vv), *_s3(gv), *_s3(ov), *_s3(dqv), code,
                      nbytes=es * (6 * B * H * N * C + 2 * B * H * N * M) + 12 * B * N * M)
                _call("clusten_scatter_rows", dev, dS.data_ptr(), qv.data_ptr(), off.data_ptr(), ent.data_ptr(), _lib.ptr(pack),
                      dkv.data_ptr(), B, H, N, N, C, M, *_s3(dS), *_s3(qv), *_s3(dkv), code,
                      nbytes=es * (B * H * N * M + 2 * B * H * N * C) + 8 * B * N * M)
                _call("clusten_scatter_rows", dev, P.data_ptr(), gv.data_ptr(), off.data_ptr(), ent.data_ptr(), _lib.ptr(pack),
                      dvv.data_ptr(), B, H, N, N, C, M, *_s3(P), *_s3(gv), *_s3(dvv), code,
                      nbytes=es * (B * H * N * M + 2 * B * H * N * C) + 8 * B * N * M)
                _call("clusten_table_grad", dev, dS.data_ptr(), bias_idx.data_ptr(), 0, d_tab.data_ptr(), B * N * M, tab.shape[0],
                      _lib.ptr(ctx.count), H,
                      N * M, H * N * M, 1, N * M, code, nbytes=es * B * H * N * M + 4 * B * N * M)
        # blank-token parameters: [H,C] column sums over all tokens -- one pass over q and d_out (clusten_blank_grad)
        if q.numel() and C % 8 == 0 and H * C <= 2048:
            d_bk = torch.zeros(H * C, dtype=torch.float32, device=dev)
            d_bv = torch.zeros(H * C, dtype=torch.float32, device=dev)
            with torch.cuda.device(dev):
                _call("clusten_blank_grad", dev, qv.data_ptr(), gv.data_ptr(), dSb.data_ptr(), Pb.data_ptr(), d_bk.data_ptr(),
                      d_bv.data_ptr(), B, H, N, C, *_s3(qv), *_s3(gv), code, nbytes=es * 2 * B * H * N * C + 8 * B * H * N)
        else:
            d_bk = torch.einsum("bhn,bnhc->hc", dSb.to(dt), q).reshape(-1)
            d_bv = torch.einsum("bhn,bnhc->hc", Pb.to(dt), d_out).reshape(-1)
        tdt, kdt, vdt = ctx.meta
        return d_q, d_kv, d_tab.to(tdt), d_bk.to(kdt), d_bv.to(vdt), None, None, None, None


def cluster_attention_core(q, kv, bias_tab, blank_k, blank_v, nbhd_idx, bias_idx, mask, count=None):
    return ClusterAttentionCoreFunction.apply(q, kv, bias_tab, blank_k, blank_v, nbhd_idx, bias_idx, mask, count)


# ---- the same core with the relative-position bias computed in the kernels (opt-in, round-2 work) -----------------------------
PE_GRAD_PARTS = 1024                      # partial-sum slots of the pos_embed gradient (posbias.cuh: PB_PARTS)


def _check_pos_args(B, N, H, pos, pe_weight, pe_bias):
    _check_shapes(pos.dtype == torch.float32 and tuple(pos.shape) == (B, N, 2), "fused attention: pos must be fp32 [B,N,2]")
    _check_shapes(tuple(pe_weight.shape) == (H, 5) and (pe_bias is None or tuple(pe_bias.shape) == (H,)),
                  "fused attention: pos_embed must be Linear(5, heads)")


def cluster_attention_fused_pos(q, key, v, nbhd_idx, pos, pe_weight, pe_bias, mask, blank_k, blank_v):
    """``cluster_attention_fused`` with the bias ``pos_embed(pre_table)[pe_idx]`` (aff.py:129-132) computed from the token
    positions inside the kernel (clusten_attn_pos_fwd): pos fp32 [B,N,2] (x, y) of the stage's tokens (queries = keys),
    pe_weight [H,5] / pe_bias [H] the parameters of ``pos_embed``.  No autograd: inference path."""
    dev = _lib.require_cuda(q, key, v, nbhd_idx, pos, pe_weight, pe_bias, mask, blank_k, blank_v)
    B, H, Nq, C = q.shape
    Nk, M = key.shape[2], nbhd_idx.shape[2]
    dt = q.dtype
    q, key, v, nbhd_idx = _rows(q), _rows(key.to(dt)), _rows(v.to(dt)), _idx(nbhd_idx)
    _check_shapes(Nq == Nk, "fused attention with in-kernel bias: queries and keys are the same token set")
    pos = pos.contiguous()
    _check_pos_args(B, Nq, H, pos, pe_weight, pe_bias)
    w = pe_weight.detach().to(torch.float32).contiguous()
    b_ = None if pe_bias is None else pe_bias.detach().to(torch.float32).contiguous()
    if mask is not None:
        _check_shapes(mask.dtype == torch.uint8 and tuple(mask.shape) == (B, Nq, M) and mask.is_contiguous(), "fused attention: mask must be uint8 [B,N,M]")
    blank_k, blank_v = blank_k.to(dt).contiguous(), blank_v.to(dt).contiguous()
    out = torch.empty((B, Nq, H, C), dtype=dt, device=dev)
    ov = out.permute(0, 2, 1, 3)
    if out.numel():
        with torch.cuda.device(dev):
            _call("clusten_attn_pos_fwd", dev, q.data_ptr(), key.data_ptr(), v.data_ptr(), nbhd_idx.data_ptr(),
                  _lib.ptr(neighbourhood_pack(nbhd_idx, Nk, mask=mask)), pos.data_ptr(), pos.data_ptr(), w.data_ptr(), _lib.ptr(b_),
                  _lib.ptr(mask), blank_k.data_ptr(), blank_v.data_ptr(), out.data_ptr(), 0, 0, B, H, Nq, Nk, C, M,
                  *_s3(q), *_s3(key), *_s3(v), *_s3(ov), _lib.dtype_code(q),
                  nbytes=q.element_size() * (B * H * (2 * Nq + 2 * Nk) * C) + 8 * B * Nq * M + 8 * B * (Nq + Nk))
    return out.reshape(B, Nq, H * C)


class ClusterAttentionPosFunction(Function):
    """``ClusterAttentionCoreFunction`` with the relative-position bias computed from positions in the kernels instead of
    gathered from ``bias_tab[bias_idx]``: no bias-index operand, no table gathers, and the gradient of ``pos_embed`` comes out
    of the backward kernel as partial sums of dS * [feat | 1] (no clusten_table_grad pass).  fp16 / bf16 training."""

    @staticmethod
    def forward(ctx, q, kv, pe_weight, pe_bias, blank_k, blank_v, nbhd_idx, pos, mask):
        dev = _lib.require_cuda(q, kv, pe_weight, pe_bias, blank_k, blank_v, nbhd_idx, pos, mask)
        _check_shapes(q.dim() == 4 and kv.dim() == 5 and kv.shape[3] == 2 and q.dtype == kv.dtype and
                      q.dtype in (torch.float16, torch.bfloat16), "fused attention: q [B,N,H,C], kv [B,N,H,2,C], fp16/bf16")
        B, N, H, C = q.shape
        M = nbhd_idx.shape[2]
        dt = q.dtype
        q, kv, nbhd_idx, pos = q.contiguous(), kv.contiguous(), _idx(nbhd_idx), pos.contiguous()
        _check_pos_args(B, N, H, pos, pe_weight, pe_bias)
        w = pe_weight.detach().to(torch.float32).contiguous()
        b_ = None if pe_bias is None else pe_bias.detach().to(torch.float32).contiguous()
        bk, bv = blank_k.detach().to(dt).contiguous(), blank_v.detach().to(dt).contiguous()
        if mask is not None:
            _check_shapes(mask.dtype == torch.uint8 and tuple(mask.shape) == (B, N, M) and mask.is_contiguous(), "fused attention: mask must be uint8 [B,N,M]")
        out = torch.empty((B, N, H, C), dtype=dt, device=dev)
        lse = torch.empty((B, H, N), dtype=torch.float32, device=dev)
        qv, kk, vv, ov = q.permute(0, 2, 1, 3), kv[:, :, :, 0].permute(0, 2, 1, 3), kv[:, :, :, 1].permute(0, 2, 1, 3), out.permute(0, 2, 1, 3)
        if out.numel():
            with torch.cuda.device(dev):
                _call("clusten_attn_pos_fwd", dev, qv.data_ptr(), kk.data_ptr(), vv.data_ptr(), nbhd_idx.data_ptr(),
                      _lib.ptr(neighbourhood_pack(nbhd_idx, N, mask=mask)), pos.data_ptr(), pos.data_ptr(), w.data_ptr(), _lib.ptr(b_),
                      _lib.ptr(mask), bk.data_ptr(), bv.data_ptr(), out.data_ptr(), 0, lse.data_ptr(), B, H, N, N, C, M,
                      *_s3(qv), *_s3(kk), *_s3(vv), *_s3(ov), _lib.dtype_code(q),
                      nbytes=q.element_size() * (B * H * 4 * N * C) + 8 * B * N * M + 16 * B * N)
        ctx.save_for_backward(q, kv, w, b_, bk, bv, out, lse, nbhd_idx, pos, mask)
        ctx.meta = (pe_weight.dtype, None if pe_bias is None else pe_bias.dtype, blank_k.dtype, blank_v.dtype)
        return out.view(B, N, H * C)

    @staticmethod
    def backward(ctx, d_out):
        q, kv, w, b_, bk, bv, out, lse, nbhd_idx, pos, mask = ctx.saved_tensors
        dev = q.device
        B, N, H, C = q.shape
        M = nbhd_idx.shape[2]
        dt = q.dtype
        d_out = d_out.to(dt).contiguous().view(B, N, H, C)
        d_q = torch.empty_like(q)
        d_kv = torch.empty_like(kv)
        P = torch.empty((B, H, N, M), dtype=dt, device=dev)
        dS = torch.empty((B, H, N, M), dtype=dt, device=dev)
        Pb = torch.empty((B, H, N), dtype=torch.float32, device=dev)
        dSb = torch.empty((B, H, N), dtype=torch.float32, device=dev)
        parts = torch.zeros((PE_GRAD_PARTS, H, 6), dtype=torch.float32, device=dev)
        hv = lambda t: t.permute(0, 2, 1, 3)
        qv, gv = hv(q), hv(d_out)
        code = _lib.dtype_code(q)
        es = q.element_size()
        if q.numel():
            kk, vv, ov, dqv = hv(kv[:, :, :, 0]), hv(kv[:, :, :, 1]), hv(out), hv(d_q)
            dkv, dvv = hv(d_kv[:, :, :, 0]), hv(d_kv[:, :, :, 1])
            pack = neighbourhood_pack(nbhd_idx, N, inverse=True, mask=mask)
            off, ent = inverse_neighbour_list(nbhd_idx, N, with_pack=True, pack_buf=pack if mask is not None else None)
            with torch.cuda.device(dev):
                _call("clusten_attn_pos_bwd", dev, gv.data_ptr(), ov.data_ptr(), lse.data_ptr(), qv.data_ptr(), kk.data_ptr(), vv.data_ptr(),
                      nbhd_idx.data_ptr(), _lib.ptr(pack), pos.data_ptr(), pos.data_ptr(), w.data_ptr(), _lib.ptr(b_), _lib.ptr(mask),
                      bk.data_ptr(), bv.data_ptr(), d_q.data_ptr(), P.data_ptr(), dS.data_ptr(), Pb.data_ptr(), dSb.data_ptr(),
                      parts.data_ptr(), B, H, N, N, C, M, *_s3(qv), *_s3(kk), *_s3(vv), *_s3(gv), *_s3(ov), *_s3(dqv), code,
                      nbytes=es * (6 * B * H * N * C + 2 * B * H * N * M) + 8 * B * N * M + 16 * B * N)
                _call("clusten_scatter_rows", dev, dS.data_ptr(), qv.data_ptr(), off.data_ptr(), ent.data_ptr(), _lib.ptr(pack),
                      dkv.data_ptr(), B, H, N, N, C, M, *_s3(dS), *_s3(qv), *_s3(dkv), code,
                      nbytes=es * (B * H * N * M + 2 * B * H * N * C) + 8 * B * N * M)
                _call("clusten_scatter_rows", dev, P.data_ptr(), gv.data_ptr(), off.data_ptr(), ent.data_ptr(), _lib.ptr(pack),
                      dvv.data_ptr(), B, H, N, N, C, M, *_s3(P), *_s3(gv), *_s3(dvv), code,
                      nbytes=es * (B * H * N * M + 2 * B * H * N * C) + 8 * B * N * M)
        if q.numel() and C % 8 == 0 and H * C <= 2048:
            d_bk = torch.zeros(H * C, dtype=torch.float32, device=dev)
            d_bv = torch.zeros(H * C, dtype=torch.float32, device=dev)
            with torch.cuda.device(dev):
                _call("clusten_blank_grad", dev, qv.data_ptr(), gv.data_ptr(), dSb.data_ptr(), Pb.data_ptr(), d_bk.data_ptr(),
                      d_bv.data_ptr(), B, H, N, C, *_s3(qv), *_s3(gv), code, nbytes=es * 2 * B * H * N * C + 8 * B * H * N)
        else:
            d_bk = torch.einsum("bhn,bnhc->hc", dSb.to(dt), q).reshape(-1)
            d_bv = torch.einsum("bhn,bnhc->hc", Pb.to(dt), d_out).reshape(-1)
        g = parts.sum(0)                                                   # [H, 6]
        wdt, bdt, kdt, vdt = ctx.meta
        d_w = g[:, :5].to(wdt)
        d_b = None if bdt is None else g[:, 5].to(bdt)
        return d_q, d_kv, d_w, d_b, d_bk.to(kdt), d_bv.to(vdt), None, None, None


def cluster_attention_core_pos(q, kv, pe_weight, pe_bias, blank_k, blank_v, nbhd_idx, pos, mask):
    return ClusterAttentionPosFunction.apply(q, kv, pe_weight, pe_bias, blank_k, blank_v, nbhd_idx, pos, mask)


# ---- LayerNorm -------------------------------------------------------------------------------------------------------
class LayerNormFunction(Function):
    """LayerNorm over the last dimension (C <= 1024), one warp per row (clusten_layer_norm_fwd / _bwd).  ``out_dtype`` lets
    the caller take the result in the dtype the consumer will cast it to anyway (the Linear layers under autocast)."""

    @staticmethod
    def forward(ctx, x, weight, bias, eps, out_dtype):
        dev = _lib.require_cuda(x, weight, bias)
        C = x.shape[-1]
        xc = x.contiguous()
        w, b = weight.detach().float().contiguous(), bias.detach().float().contiguous()
        R = xc.numel() // C
        y = torch.empty(xc.shape, dtype=out_dtype, device=dev)
        need = any(ctx.needs_input_grad[:3])          # (grad mode is off inside Function.forward: ask autograd instead)
        mean = torch.empty(R, dtype=torch.float32, device=dev) if need else None
        rstd = torch.empty(R, dtype=torch.float32, device=dev) if need else None
        if R:
            with torch.cuda.device(dev):
                _call("clusten_layer_norm_fwd", dev, xc.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), _lib.ptr(mean),
                      _lib.ptr(rstd), R, C, float(eps), _lib.dtype_code(xc), _lib.DTYPES[out_dtype],
                      nbytes=R * C * (xc.element_size() + y.element_size()))
        ctx.save_for_backward(xc, w, mean, rstd)
        ctx.wdt = (weight.dtype, bias.dtype)
        return y

    @staticmethod
    def backward(ctx, dy):
        xc, w, mean, rstd = ctx.saved_tensors
        dev = xc.device
        C = xc.shape[-1]
        R = xc.numel() // C
        dy = dy.contiguous()
        if dy.dtype not in _lib.DTYPES:
            dy = dy.float()
        dx = torch.empty_like(xc)
        dg = torch.zeros(C, dtype=torch.float32, device=dev)
        db = torch.zeros(C, dtype=torch.float32, device=dev)
        if R:
            with torch.cuda.device(dev):
                _call("clusten_layer_norm_bwd", dev, dy.data_ptr(), xc.data_ptr(), w.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                      dx.data_ptr(), dg.data_ptr(), db.data_ptr(), R, C, _lib.dtype_code(xc), _lib.dtype_code(dy),
                      nbytes=R * C * (2 * xc.element_size() + dy.element_size()))
        return dx, dg.to(ctx.wdt[0]), db.to(ctx.wdt[1]), None, None


def layer_norm_stats(x, weight, bias, eps=1e-5):
    """(mean, rstd) fp32 [rows] of LayerNorm over the last dimension, exactly the statistics clusten_layer_norm_fwd normalises with
    (same kernel, y = NULL): what ``linear_tc(..., ln=...)`` needs to apply the norm while it stages the rows.  No autograd."""
    dev = _lib.require_cuda(x, weight, bias)
    C = x.shape[-1]
    xc = x.contiguous()
    R = xc.numel() // C
    mean = torch.empty(R, dtype=torch.float32, device=dev)
    rstd = torch.empty(R, dtype=torch.float32, device=dev)
    if R:
        w, b = weight.detach().float().contiguous(), bias.detach().float().contiguous()
        with torch.cuda.device(dev):
            _call("clusten_layer_norm_fwd", dev, xc.data_ptr(), w.data_ptr(), b.data_ptr(), 0, mean.data_ptr(), rstd.data_ptr(), R, C,
                  float(eps), _lib.dtype_code(xc), _lib.DTYPES[torch.float32], nbytes=R * C * xc.element_size() + 8 * R)
    return mean, rstd


def layer_norm(x, weight, bias, eps=1e-5, out_dtype=None):
    return LayerNormFunction.apply(x, weight, bias, eps, out_dtype or x.dtype)


# ---- residual + layer scale + stochastic depth -----------------------------------------------------------------------
def _residual_out_dtype(res, x, gamma):
    """Type promotion of ``res + x * gamma * scale`` as ATen does it (gamma is a dimensioned tensor: fp32 gamma lifts 16-bit x)."""
    dt = torch.promote_types(res.dtype, x.dtype)
    return dt if gamma is None else torch.promote_types(dt, gamma.dtype)


def scale_residual_supported(res, x, gamma, sample_scale):
    if not (res.is_cuda and res.dim() >= 2 and res.shape == x.shape and res.is_contiguous() and x.is_contiguous()):
        return False
    C = res.shape[-1]
    odt = _residual_out_dtype(res, x, gamma)
    f32 = torch.float32
    combo = (res.dtype, x.dtype, odt)
    ok = combo == (f32, f32, f32) or (x.dtype in (torch.float16, torch.bfloat16) and (
        combo == (f32, x.dtype, f32) or combo == (x.dtype, x.dtype, x.dtype) or combo == (x.dtype, x.dtype, f32)))
    if gamma is not None and (gamma.dtype != f32 or gamma.shape != (C,)):
        return False
    if sample_scale is not None and (sample_scale.dtype != f32 or sample_scale.shape != (res.shape[0],)):
        return False
    return bool(ok and C % 4 == 0 and C <= 1024 and res.shape[0] <= 65535 and res.numel() > 0
                and res.data_ptr() % 16 == 0 and x.data_ptr() % 16 == 0)


class ScaleResidualFunction(Function):
    """``res + x * gamma * sample_scale`` in one pass (clusten_scale_residual_fwd / _bwd): the residual lines of the block,
    backbone/aff.py:230,236, with the layer scale (gamma [C] or None) and the per-sample stochastic-depth factor
    (sample_scale [B] = bernoulli(keep) / keep or None) folded in.  d_gamma accumulates with fp32 atomics."""

    @staticmethod
    def forward(ctx, res, x, gamma, sample_scale):
        dev = _lib.require_cuda(res, x)
        B, C = res.shape[0], res.shape[-1]
        rows = res.numel() // (B * C)
        odt = _residual_out_dtype(res, x, gamma)
        out = torch.empty(res.shape, dtype=odt, device=dev)
        g = None if gamma is None else gamma.detach().contiguous()
        with torch.cuda.device(dev):
            _call("clusten_scale_residual_fwd", dev, res.data_ptr(), x.data_ptr(), _lib.ptr(g), _lib.ptr(sample_scale),
                  out.data_ptr(), B, rows, C, _lib.dtype_code(res), _lib.dtype_code(x), _lib.DTYPES[odt],
                  nbytes=res.numel() * (res.element_size() + x.element_size() + out.element_size()))
        need_dgamma = gamma is not None and ctx.needs_input_grad[2]
        ctx.save_for_backward(x if need_dgamma else None, g, sample_scale)
        ctx.meta = (res.dtype, x.dtype, B, rows, C)
        return out

    @staticmethod
    def backward(ctx, d_out):
        x, g, s = ctx.saved_tensors
        rdt, xdt, B, rows, C = ctx.meta
        dev = d_out.device
        d_out = d_out.contiguous()
        d_res = d_x = d_gamma = None
        if ctx.needs_input_grad[0]:
            d_res = d_out if d_out.dtype == rdt else d_out.to(rdt)
        want_x, want_g = ctx.needs_input_grad[1], x is not None
        if want_x or want_g:
            if g is None and s is None and want_x:
                d_x = d_out if d_out.dtype == xdt else d_out.to(xdt)
            else:
                if want_x:
                    d_x = torch.empty(d_out.shape, dtype=xdt, device=dev)
                if want_g:
                    d_gamma = torch.zeros(C, dtype=torch.float32, device=dev)
                with torch.cuda.device(dev):
                    _call("clusten_scale_residual_bwd", dev, d_out.data_ptr(), _lib.ptr(x), _lib.ptr(g), _lib.ptr(s), _lib.ptr(d_x),
                          _lib.ptr(d_gamma), B, rows, C, _lib.dtype_code(d_out), _lib.DTYPES[xdt],
                          nbytes=d_out.numel() * (d_out.element_size() + (2 if want_g else 1) * (4 if xdt == torch.float32 else 2)))
        return d_res, d_x, d_gamma, None


def scale_residual(res, x, gamma=None, sample_scale=None):
    """``res + x * gamma[c] * sample_scale[b]``; shapes / dtypes the kernel does not take go through the op-by-op formulation."""
    if gamma is None and sample_scale is None:
        return res + x
    if scale_residual_supported(res, x, gamma, sample_scale):
        return ScaleResidualFunction.apply(res, x, gamma, sample_scale)
    if gamma is not None:
        x = gamma * x
    if sample_scale is not None:
        x = x * sample_scale.to(x.dtype).view(-1, *([1] * (x.dim() - 1)))
    return res + x


# ---- Linear over the relative-position feature table ------------------------------------------------------------------
class TableLinearFunction(Function):
    """``F.linear(features, weight, bias)`` for the [R, F <= 8] relative-position feature table and H <= 32 outputs
    (pos_embed = Linear(5, heads), aff.py:101,129) in fp32, restricted to the first ``count`` rows (device int32 scalar, or
    None = all): rows past it come back as zeros and take no part in the weight gradient.  features get no gradient."""

    @staticmethod
    def forward(ctx, features, weight, bias, count):
        dev = _lib.require_cuda(features, weight, bias, count)
        R, F = features.shape
        H = weight.shape[0]
        feat = features.detach().contiguous()
        w = weight.detach().contiguous()
        b = None if bias is None else bias.detach().contiguous()
        out = torch.empty((R, H), dtype=torch.float32, device=dev)
        if R:
            with torch.cuda.device(dev):
                _call("clusten_table_linear_fwd", dev, feat.data_ptr(), w.data_ptr(), _lib.ptr(b), out.data_ptr(), R, F, H,
                      _lib.ptr(count), nbytes=4 * R * (F + H))
        ctx.save_for_backward(feat, count)
        ctx.meta = (R, F, H, weight.dtype, None if bias is None else bias.dtype)
        return out

    @staticmethod
    def backward(ctx, g):
        feat, count = ctx.saved_tensors
        R, F, H, wdt, bdt = ctx.meta
        dev = g.device
        g = g.to(torch.float32).contiguous()
        dW = torch.zeros((H, F), dtype=torch.float32, device=dev)
        db = torch.zeros(H, dtype=torch.float32, device=dev) if bdt is not None else None
        if R:
            with torch.cuda.device(dev):
                _call("clusten_table_linear_bwd", dev, g.data_ptr(), feat.data_ptr(), dW.data_ptr(), _lib.ptr(db), R, F, H,
                      _lib.ptr(count), nbytes=4 * R * (F + H))
        return None, dW.to(wdt), None if db is None else db.to(bdt), None


def table_linear_supported(features, weight, bias):
    return bool(features.is_cuda and features.dim() == 2 and features.dtype == torch.float32 and weight.dtype == torch.float32
                and (bias is None or bias.dtype == torch.float32) and features.shape[1] == weight.shape[1] <= 8
                and weight.shape[0] <= 32 and features.shape[0] * weight.shape[0] < 2 ** 31)


def table_linear(features, weight, bias=None, count=None):
    return TableLinearFunction.apply(features, weight, bias, count)


# ---- Linear with a bandwidth-bound bias gradient -----------------------------------------------------------------------
def col_sum(x2d):
    """fp32 [C] column sums of a [R, C] matrix (clusten_col_sum); falls back to torch for shapes the kernel does not take."""
    R, C = x2d.shape
    vpt = 4 if x2d.dtype == torch.float32 else 8
    if (x2d.dtype not in _lib.DTYPES or C % vpt or x2d.stride(1) != 1 or x2d.stride(0) % vpt or x2d.data_ptr() % 16 or R == 0
            or R >= 2 ** 31):
        return x2d.float().sum(0)
    out = torch.zeros(C, dtype=torch.float32, device=x2d.device)
    with torch.cuda.device(x2d.device):
        _call("clusten_col_sum", x2d.device, x2d.data_ptr(), out.data_ptr(), R, C, x2d.stride(0), _lib.dtype_code(x2d),
              nbytes=x2d.element_size() * R * C)
    return out


def linear_f32_supported(x, weight, bias):
    K = x.shape[-1]
    return bool(x.is_cuda and x.dtype == torch.float32 and weight.dtype == torch.float32 and weight.dim() == 2 and weight.shape[1] == K
                and (bias is None or bias.dtype == torch.float32) and K % 32 == 0 and weight.shape[0] % 2 == 0 and x.numel() > 0
                and x.numel() // K < 2 ** 31 - 128)


def linear_f32(x, weight, bias=None):
    """fp32 ``F.linear`` on the tensor cores with the 3xTF32 split (clusten_linear_f32; inference, no autograd).  Opt-in: the
    caller checks ``linear_f32_supported`` first."""
    dev = _lib.require_cuda(x, weight, bias)
    K, N = x.shape[-1], weight.shape[0]
    x2 = x.reshape(-1, K)
    if x2.stride(1) != 1 or x2.stride(0) % 4 or x2.data_ptr() % 16:
        x2 = x2.contiguous()
    w = weight.detach().contiguous()
    b = None if bias is None else bias.detach().contiguous()
    y = torch.empty((x2.shape[0], N), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _call("clusten_linear_f32", dev, x2.data_ptr(), w.data_ptr(), _lib.ptr(b), y.data_ptr(), x2.shape[0], K, N, x2.stride(0), N,
              nbytes=4 * (x2.shape[0] * (K + N) + N * K))
    return y.view(*x.shape[:-1], N)


# ---- fp32 Linear on tcgen05 (3xTF32) with the following element-wise line folded in ----------------------------------------------
LINEAR_EPI = {"bias": 0, "gelu": 1, "residual": 2}
_split_cache = {}                                      # id(weight) -> (weakref, version, data_ptr, hi, lo)
LINEAR_TC_CHAIN = int(os.environ.get("CLUSTEN_TC_CHAIN", "0"))


def tf32_split(weight):
    """(hi, lo) of an fp32 weight for clusten_linear_tc_f32 (clusten_tf32_split), cached until the weight changes."""
    import weakref
    key = id(weight)
    hit = _split_cache.get(key)
    if hit is not None and hit[0]() is weight and hit[1] == weight._version and hit[2] == weight.data_ptr():
        return hit[3], hit[4]
    dev = _lib.require_cuda(weight)
    w = weight.detach().contiguous()
    hi, lo = torch.empty_like(w), torch.empty_like(w)
    with torch.cuda.device(dev):
        _call("clusten_tf32_split", dev, w.data_ptr(), hi.data_ptr(), lo.data_ptr(), w.numel(), nbytes=12 * w.numel())
    if len(_split_cache) > 4096:
        _split_cache.clear()
    _split_cache[key] = (weakref.ref(weight), weight._version, weight.data_ptr(), hi, lo)
    return hi, lo


# which split the tcgen05 Linear multiplies with: "f16" (fp16 hi / lo, kind::f16: half the MMAs) or "tf32" (TF32 hi / lo, kind::tf32)
# Default ("auto"): fp16 where the input scale is known -- the LayerNorm-fused layers, whose rows are normalised -- and TF32 elsewhere
# (fp16 has no exponent headroom for x: elements below 2^-14 keep an absolute error of 2^-25, values beyond 65504 saturate).
LINEAR_TC_SPLIT = os.environ.get("CLUSTEN_TC_SPLIT", "auto")
_split16_cache = {}


def f16_split(weight):
    """(hi, lo, inv_scale) of an fp32 weight [N, K] for clusten_linear_tc_f32(w_fp16 = 1) (clusten_f16_split): fp16 halves of
    w[n] * s_n with s_n the power of two that brings max |w[n]| to [512, 1024), and 1 / s_n as a device vector [N].  No host read;
    cached until the weight changes."""
    import weakref
    key = id(weight)
    hit = _split16_cache.get(key)
    if hit is not None and hit[0]() is weight and hit[1] == weight._version and hit[2] == weight.data_ptr():
        return hit[3], hit[4], hit[5]
    dev = _lib.require_cuda(weight)
    w = weight.detach().contiguous()
    hi = torch.empty(w.shape, dtype=torch.float16, device=dev)
    lo = torch.empty_like(hi)
    amax = w.abs().amax(dim=1).float().contiguous()
    inv = torch.empty(w.shape[0], dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _call("clusten_f16_split", dev, w.data_ptr(), hi.data_ptr(), lo.data_ptr(), w.shape[0], w.shape[1], amax.data_ptr(), inv.data_ptr(),
              nbytes=8 * w.numel())
    if len(_split16_cache) > 4096:
        _split16_cache.clear()
    _split16_cache[key] = (weakref.ref(weight), weight._version, weight.data_ptr(), hi, lo, inv)
    return hi, lo, inv


# LayerNorm in front of the layer: gamma / beta folded into the layer (W' = W * gamma per input column, b' = b + W beta; computed once per
# parameter version in float64) so that the GEMM's split warps only normalise, (x - mean) * rstd.  CLUSTEN_LN_FOLD=0: the affine part
# is applied while the rows are split (bit-identical to LayerNorm-then-Linear; ~90 more instructions per K chunk in the critical role).
LN_FOLD = os.environ.get("CLUSTEN_LN_FOLD", "1") != "0"
_lnfold_cache = {}                                     # id(weight) -> (weakrefs, versions, W', b')


def ln_folded_layer(weight, bias, ln_weight, ln_bias):
    """(W', b') with ``F.linear(F.layer_norm(x, w=ln_weight, b=ln_bias), weight, bias) == F.linear(normalise(x), W', b')``; cached until
    one of the four parameters changes (version counters / storage)."""
    import weakref
    ts = (weight, bias, ln_weight, ln_bias)
    sig = tuple((None if t is None else (t._version, t.data_ptr())) for t in ts)
    hit = _lnfold_cache.get(id(weight))
    if hit is not None and hit[1] == sig and all((r is None and t is None) or (r is not None and r() is t) for r, t in zip(hit[0], ts)):
        return hit[2], hit[3]
    with torch.no_grad():
        w64 = weight.detach().double()
        wf = (w64 * ln_weight.detach().double().unsqueeze(0)).float().contiguous()
        bf = w64 @ ln_bias.detach().double()
        if bias is not None:
            bf = bf + bias.detach().double()
        bf = bf.float().contiguous()
    if len(_lnfold_cache) > 4096:
        _lnfold_cache.clear()
    _lnfold_cache[id(weight)] = (tuple(None if t is None else weakref.ref(t) for t in ts), sig, wf, bf)
    return wf, bf


def linear_tc_supported(x, weight, bias=None, res=None, gamma=None):
    K = x.shape[-1]
    ok = (x.is_cuda and x.dtype == torch.float32 and weight.dtype == torch.float32 and weight.dim() == 2 and weight.shape[1] == K
          and K % 32 == 0 and weight.shape[0] % 4 == 0 and x.numel() > 0 and x.numel() // K < 2 ** 31 - 128)
    for t in (bias, res, gamma):
        ok = ok and (t is None or (t.dtype == torch.float32 and t.is_cuda))
    return bool(ok)


def linear_tc(x, weight, bias=None, epilogue="bias", res=None, gamma=None, alpha=1.0, alpha_cols=0, chain=None, ln=None, split=None,
              ln_fold=None):
    """fp32 ``F.linear`` on the tcgen05 tensor cores (clusten_linear_tc_f32; inference, no autograd) with one of
    ``bias`` (y = x W^T + b, the first ``alpha_cols`` columns then times ``alpha``), ``gelu`` (y = GELU(x W^T + b)) or
    ``residual`` (y = res + gamma * (x W^T + b)) as the epilogue.  ``ln`` = (mean, rstd, ln_weight, ln_bias): the rows of x are
    LayerNorm-ed on the fly with the statistics of ``layer_norm_stats``; ``ln_fold`` (default: LN_FOLD) folds its gamma / beta into the
    weights and the bias (``ln_folded_layer``) so that the kernel only normalises.  The caller checks ``linear_tc_supported`` first."""
    dev = _lib.require_cuda(x, weight, bias, res, gamma)
    K, N = x.shape[-1], weight.shape[0]
    fold = ln is not None and (LN_FOLD if ln_fold is None else bool(ln_fold))
    if fold:
        _check_shapes(ln[2].numel() == K and ln[3].numel() == K, "linear_tc: LayerNorm operands")
        weight, bias = ln_folded_layer(weight, bias, ln[2], ln[3])
    x2 = x.reshape(-1, K)
    if x2.stride(1) != 1 or x2.stride(0) % 4 or x2.data_ptr() % 16:
        x2 = x2.contiguous()
    mode = split or LINEAR_TC_SPLIT
    half = mode == "f16" or (mode == "auto" and ln is not None)
    if half:
        hi, lo, inv = f16_split(weight)
    else:
        (hi, lo), inv = tf32_split(weight), None
    b = None if bias is None else bias.detach().contiguous()
    g = None if gamma is None else gamma.detach().contiguous()
    R = x2.shape[0]
    r2, ldres = None, 0
    if epilogue == "residual":
        r2 = res.reshape(-1, N)
        if r2.stride(1) != 1 or r2.stride(0) % 4 or r2.data_ptr() % 16:
            r2 = r2.contiguous()
        ldres = r2.stride(0)
    y = torch.empty((R, N), dtype=torch.float32, device=dev)
    lnp = [0, 0, 0, 0]
    if ln is not None:
        mean, rstd, lw, lb = ln
        lw, lb = lw.detach().float().contiguous(), lb.detach().float().contiguous()
        _check_shapes(mean.numel() == R and rstd.numel() == R and lw.numel() == K and lb.numel() == K, "linear_tc: LayerNorm operands")
        lnp = [mean.data_ptr(), rstd.data_ptr(), 0, 0] if fold else [mean.data_ptr(), rstd.data_ptr(), lw.data_ptr(), lb.data_ptr()]
    with torch.cuda.device(dev):
        _call("clusten_linear_tc_f32", dev, x2.data_ptr(), hi.data_ptr(), lo.data_ptr(), _lib.ptr(b), _lib.ptr(r2), _lib.ptr(g),
              y.data_ptr(), R, K, N, x2.stride(0), N, ldres, LINEAR_EPI[epilogue], float(alpha), int(alpha_cols),
              LINEAR_TC_CHAIN if chain is None else int(chain), *lnp, int(half), _lib.ptr(inv), nbytes=4 * (R * (K + N * (2 if r2 is not None else 1)) + 2 * N * K),
              flops=(2 * R * K * N * (3 if half else 6), 2 * R * K * N))
    return y.view(*x.shape[:-1], N)


class LinearFunction(Function):
    """``F.linear`` whose backward takes the bias gradient with clusten_col_sum (one coalesced pass, fp32 accumulation) instead
    of ATen's generic reduction; the two GEMMs of the backward stay in cuBLAS.  Autocast-aware like the native op: under
    autocast the operands are cast once, the casts are what is saved, and the gradients come back in the parameters' dtypes."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        if torch.is_autocast_enabled():
            dt = torch.get_autocast_dtype("cuda")
            xc, wc = x.to(dt), weight.to(dt)
            bc = None if bias is None else bias.to(dt)
        else:
            xc, wc, bc = x, weight, bias
        with torch.autocast("cuda", enabled=False):
            y = torch.nn.functional.linear(xc, wc, bc)
        ctx.save_for_backward(xc, wc)
        ctx.meta = (x.dtype, weight.dtype, None if bias is None else bias.dtype)
        return y

    @staticmethod
    def backward(ctx, gy):
        xc, wc = ctx.saved_tensors
        xdt, wdt, bdt = ctx.meta
        g2 = gy.reshape(-1, gy.shape[-1])
        if g2.dtype != wc.dtype:
            g2 = g2.to(wc.dtype)
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            gx = (g2 @ wc).view(xc.shape).to(xdt)
        if ctx.needs_input_grad[1]:
            gw = (g2.t() @ xc.reshape(-1, xc.shape[-1])).to(wdt)
        if bdt is not None and ctx.needs_input_grad[2]:
            gb = col_sum(g2 if g2.is_contiguous() else g2.contiguous()).to(bdt)
        return gx, gw, gb


def linear(x, weight, bias=None):
    return LinearFunction.apply(x, weight, bias)


# ---- relative-position table lookup ----------------------------------------------------------------------------------
class TableLookupFunction(Function):
    """``tab[inverse]`` with a backward that does not go through ATen's sort-based ``index_put_(accumulate=True)``
    (hundreds of ms for the ~25 M duplicated indices of an AFF stage): out[..., c] = tab[inverse[...], c].

    tab [U, CH] (fp32 / fp16 / bf16), inverse int64 or int32 of any shape -> out [*inverse.shape, CH].  The gradient is a
    shared-memory-privatised fp32 segment sum (clusten_table_grad); its summation order is not fixed (fp32 atomics)."""

    @staticmethod
    def forward(ctx, tab, inverse, count=None):
        dev = _lib.require_cuda(tab, inverse)
        _check_shapes(tab.dim() == 2 and inverse.dtype in (torch.int64, torch.int32), "table lookup: tab [U,CH], inverse int32/int64")
        tab = tab.contiguous()
        inverse = inverse.contiguous()
        U, CH = tab.shape
        out = torch.empty((*inverse.shape, CH), dtype=tab.dtype, device=dev)
        n = inverse.numel()
        if n:
            with torch.cuda.device(dev):
                _call("clusten_table_gather", dev, tab.data_ptr(), inverse.data_ptr(), int(inverse.dtype == torch.int64),
                      out.data_ptr(), n, U, CH, _lib.dtype_code(tab))
        ctx.save_for_backward(inverse)
        ctx.count = count                  # device int32 scalar: rows actually referenced (tab may be sized by an upper bound)
        ctx.tab_shape, ctx.tab_dtype = (U, CH), tab.dtype
        return out

    @staticmethod
    def backward(ctx, grad):
        (inverse,) = ctx.saved_tensors
        U, CH = ctx.tab_shape
        dev = grad.device
        d_tab = torch.zeros((U, CH), dtype=torch.float32, device=dev)
        n = inverse.numel()
        if n:
            if grad.dtype not in _lib.DTYPES:
                grad = grad.float()
            # address grad as base + b*sb + e*se + c*sc with e running over all but the first index dimension
            g = grad.reshape(inverse.shape[0], -1, CH) if inverse.dim() > 1 else grad.reshape(1, -1, CH)
            if g.data_ptr() != grad.data_ptr():              # reshape had to copy: now contiguous
                g = g.contiguous()
            with torch.cuda.device(dev):
                _call("clusten_table_grad", dev, g.data_ptr(), inverse.data_ptr(), int(inverse.dtype == torch.int64),
                      d_tab.data_ptr(), n, U, _lib.ptr(ctx.count), CH, g.shape[1], g.stride(0), g.stride(1), g.stride(2), _lib.dtype_code(g))
        return d_tab.to(ctx.tab_dtype), None, None


def table_lookup(tab, inverse, count=None):
    return TableLookupFunction.apply(tab, inverse, count)


# ---- WF --------------------------------------------------------------------------------------------------------------
class CLUSTENWFFunction(Function):
    """weights times feature: feat_new[b,i,ic,c] = sum_j weights[b,i,j,ic] * feat[b,nbhd_idx[b,i,j],c]  (clusten.py:71-94)"""

    @staticmethod
    def forward(ctx, weights, feat, nbhd_idx):
        dev = _lib.require_cuda(weights, feat, nbhd_idx)
        if feat.dtype != weights.dtype:
            feat = feat.to(weights.dtype)                               # clusten.py:80-81
        _check_shapes(weights.dim() == 4 and feat.dim() == 3 and nbhd_idx.dim() == 3, "WF: weights 4-D, feat/nbhd_idx 3-D")
        B, Nq, M, IC = weights.shape
        Nk, C = feat.shape[1], feat.shape[2]
        _check_shapes(feat.shape[0] == B and tuple(nbhd_idx.shape) == (B, Nq, M), "WF: shape mismatch")
        weights, feat, nbhd_idx = weights.contiguous(), _rows(feat), _idx(nbhd_idx)
        out = torch.empty((B, Nq, IC, C), dtype=weights.dtype, device=dev)
        if out.numel():
            plan = wf_plan(nbhd_idx, Nk) if weights.element_size() == 2 else None
            with torch.cuda.device(dev):
                _call("clusten_wf_fwd", dev, weights.data_ptr(), feat.data_ptr(), nbhd_idx.data_ptr(), _lib.ptr(plan), out.data_ptr(),
                      B, Nq, Nk, C, M, IC, feat.stride(0), feat.stride(1), _lib.dtype_code(weights),
                      nbytes=weights.element_size() * (B * Nq * M * IC + B * Nk * C + B * Nq * IC * C) + 8 * B * Nq * M)
        ctx.save_for_backward(weights, feat, nbhd_idx)
        return out

    @staticmethod
    def backward(ctx, grad_feat_new):
        weights, feat, nbhd_idx = ctx.saved_tensors
        dev = weights.device
        B, Nq, M, IC = weights.shape
        Nk, C = feat.shape[1], feat.shape[2]
        grad_feat_new = grad_feat_new.contiguous()
        if grad_feat_new.dtype != weights.dtype:
            grad_feat_new = grad_feat_new.to(weights.dtype)
        d_weights = torch.empty_like(weights)
        d_feat = torch.empty((B, Nk, C), dtype=weights.dtype, device=dev)
        if Nq * M == 0 or B == 0:
            return d_weights, d_feat.zero_(), None
        # the octet-form d_f kernels (16-bit tensor-core form, fp32 SIMT form) need the plan; fp32 is the AMP merge (clusten.py:80-81)
        es = weights.element_size()
        plan = wf_plan(nbhd_idx, Nk) if IC == 4 and ((es == 2 and C % 16 == 0) or (es == 4 and C % 32 == 0)) else None
        off, ent = inverse_neighbour_list(nbhd_idx, Nk, wf_plan_buf=plan)
        with torch.cuda.device(dev):
            _call("clusten_wf_bwd", dev, grad_feat_new.data_ptr(), weights.data_ptr(), feat.data_ptr(),
                  nbhd_idx.data_ptr(), off.data_ptr(), ent.data_ptr(), _lib.ptr(plan), d_weights.data_ptr(), d_feat.data_ptr(),
                  B, Nq, Nk, C, M, IC, feat.stride(0), feat.stride(1), d_feat.stride(0), d_feat.stride(1),
                  _lib.dtype_code(weights),
                  nbytes=weights.element_size() * (B * Nq * IC * C + 2 * B * Nq * M * IC + 2 * B * Nk * C) + 8 * B * Nq * M)
        return d_weights, d_feat, None


# ---- WEIGHTEDGATHER --------------------------------------------------------------------------------------------------
class WEIGHTEDGATHERFunction(Function):
    """weighted gather: feat_new[b,i,c] = sum_k weights[b,i,k] * feat[b,nbhd_idx[b,i,k],c]   (clusten.py:97-120)"""

    @staticmethod
    def forward(ctx, nbhd_idx, weights, feat):
        dev = _lib.require_cuda(nbhd_idx, weights, feat)
        if feat.dtype != weights.dtype:
            weights = weights.to(feat.dtype)                            # clusten.py:106-107
        _check_shapes(weights.dim() == 3 and feat.dim() == 3 and nbhd_idx.dim() == 3, "WG: all operands 3-D")
        B, Nq, K = weights.shape
        Nk, C = feat.shape[1], feat.shape[2]
        _check_shapes(feat.shape[0] == B and tuple(nbhd_idx.shape) == (B, Nq, K), "WG: shape mismatch")
        weights, feat, nbhd_idx = weights.contiguous(), _rows(feat), _idx(nbhd_idx)
        out = torch.empty((B, Nq, C), dtype=feat.dtype, device=dev)
        if out.numel():
            with torch.cuda.device(dev):
                _call("clusten_wg_fwd", dev, nbhd_idx.data_ptr(), weights.data_ptr(), feat.data_ptr(), out.data_ptr(),
                      B, Nq, Nk, C, K, feat.stride(0), feat.stride(1), _lib.dtype_code(feat),
                      nbytes=feat.element_size() * (B * Nq * K + B * Nk * C + B * Nq * C) + 8 * B * Nq * K)
        ctx.save_for_backward(nbhd_idx, weights, feat)
        return out

    @staticmethod
    def backward(ctx, grad_feat_new):
        nbhd_idx, weights, feat = ctx.saved_tensors
        dev = feat.device
        B, Nq, K = weights.shape
        Nk, C = feat.shape[1], feat.shape[2]
        grad_feat_new = grad_feat_new.contiguous()
        if grad_feat_new.dtype != feat.dtype:
            grad_feat_new = grad_feat_new.to(feat.dtype)
        d_weights = torch.empty_like(weights)
        d_feat = torch.empty((B, Nk, C), dtype=feat.dtype, device=dev)
        if Nq * K == 0 or B == 0:
            return None, d_weights, d_feat.zero_()
        off, ent = inverse_neighbour_list(nbhd_idx, Nk)
        with torch.cuda.device(dev):
            _call("clusten_wg_bwd", dev, grad_feat_new.data_ptr(), nbhd_idx.data_ptr(), weights.data_ptr(),
                  feat.data_ptr(), off.data_ptr(), ent.data_ptr(), d_weights.data_ptr(), d_feat.data_ptr(),
                  B, Nq, Nk, C, K, feat.stride(0), feat.stride(1), d_feat.stride(0), d_feat.stride(1),
                  _lib.dtype_code(feat),
                  nbytes=feat.element_size() * (B * Nq * C + 2 * B * Nq * K + 2 * B * Nk * C) + 8 * B * Nq * K)
        return None, d_weights, d_feat


# ---- MSDETRPC --------------------------------------------------------------------------------------------------------
class MSDETRPCFunction(Function):
    """deformable multi-scale DETR attention on point clouds (clusten.py:123-146, msdetrpc_cuda_kernel.cu:18-55):

        feat[b,i,c] = sum_m attn[b,i,m] * sum_k nn_weight[b,i,m,k] * val[b, nn_idx[b,i,m,k], c]

    One kernel forward (clusten_msdetrpc_fwd: the weighted-gather kernel with the product weight attn[m] * nn_weight[m,k] formed in
    shared memory), two backward (clusten_msdetrpc_bwd: d_nn_weight and d_attn from the product's gradient in one pass;
    d_val by the deterministic inverse-list gather instead of the reference's atomics, msdetrpc_cuda_kernel.cu:113-131).
    ``.apply(nn_idx, nn_weight, attn, val)`` -> [B,N,C]; backward returns (None, d_weight, d_attn, d_val)."""

    @staticmethod
    def forward(ctx, nn_idx, nn_weight, attn, val):
        dev = _lib.require_cuda(nn_idx, nn_weight, attn, val)
        _check_shapes(nn_idx.dim() == 4 and nn_weight.shape == nn_idx.shape and attn.dim() == 3 and val.dim() == 3 and
                      tuple(attn.shape) == tuple(nn_idx.shape[:3]) and val.shape[0] == nn_idx.shape[0], "MSDETRPC: shape mismatch")
        if nn_idx.dtype != torch.int64:
            raise RuntimeError(f"nn_idx must be int64 (got {nn_idx.dtype})")
        B, N, M, K = nn_idx.shape
        Nk, C = val.shape[1], val.shape[2]
        dt = val.dtype
        nn_idx, val = nn_idx.contiguous(), _rows(val)
        ctx.in_dtypes = (nn_weight.dtype, attn.dtype)
        nn_weight, attn = nn_weight.contiguous().to(dt), attn.contiguous().to(dt)
        out = torch.empty((B, N, C), dtype=dt, device=dev)
        ctx.vector = C % (16 // val.element_size()) == 0 and val.data_ptr() % 16 == 0 and (val.stride(0) * val.element_size()) % 16 == 0 \
            and (val.stride(1) * val.element_size()) % 16 == 0
        if out.numel() and not ctx.vector:
            # channel counts off the 16-byte grid: the scalar weighted-gather kernels with the product weights formed by torch
            w = (attn.unsqueeze(3) * nn_weight).reshape(B, N, M * K)
            with torch.cuda.device(dev):
                _call("clusten_wg_fwd", dev, nn_idx.data_ptr(), w.data_ptr(), val.data_ptr(), out.data_ptr(),
                      B, N, Nk, C, M * K, val.stride(0), val.stride(1), _lib.dtype_code(val),
                      nbytes=val.element_size() * (2 * B * N * M * K + B * Nk * C + B * N * C) + 8 * B * N * M * K)
        elif out.numel():
            with torch.cuda.device(dev):
                _call("clusten_msdetrpc_fwd", dev, nn_idx.data_ptr(), nn_weight.data_ptr(), attn.data_ptr(), val.data_ptr(), out.data_ptr(),
                      B, N, Nk, C, M, K, val.stride(0), val.stride(1), _lib.dtype_code(val),
                      nbytes=val.element_size() * (B * N * M * K + B * N * M + B * Nk * C + B * N * C) + 8 * B * N * M * K)
        ctx.save_for_backward(nn_idx, nn_weight, attn, val)
        return out

    @staticmethod
    def backward(ctx, grad_feat):
        nn_idx, nn_weight, attn, val = ctx.saved_tensors
        dev = val.device
        B, N, M, K = nn_idx.shape
        Nk, C = val.shape[1], val.shape[2]
        grad_feat = grad_feat.contiguous().to(val.dtype)
        d_val = torch.empty((B, Nk, C), dtype=val.dtype, device=dev)
        d_weight, d_attn = torch.empty_like(nn_weight), torch.empty_like(attn)
        wdt, adt = ctx.in_dtypes
        if B * N * M * K == 0:
            return None, d_weight.zero_().to(wdt), d_attn.zero_().to(adt), d_val.zero_()
        idx3 = nn_idx.view(B, N, M * K)
        cached = getattr(nn_idx, "_clusten_csr", None)           # built by an earlier backward on the same index tensor
        if cached is not None:
            idx3._clusten_csr = cached
        off, ent = inverse_neighbour_list(idx3, Nk)
        try:                                                     # keep the list cached on the tensor the caller holds
            nn_idx._clusten_csr = idx3._clusten_csr
        except Exception:  # pragma: no cover
            pass
        if not ctx.vector:
            w = (attn.unsqueeze(3) * nn_weight).reshape(B, N, M * K)
            d_w = torch.empty_like(w)
            with torch.cuda.device(dev):
                _call("clusten_wg_bwd", dev, grad_feat.data_ptr(), idx3.data_ptr(), w.data_ptr(), val.data_ptr(), off.data_ptr(),
                      ent.data_ptr(), d_w.data_ptr(), d_val.data_ptr(), B, N, Nk, C, M * K, val.stride(0), val.stride(1),
                      d_val.stride(0), d_val.stride(1), _lib.dtype_code(val),
                      nbytes=val.element_size() * (B * N * C + 2 * B * N * M * K + 2 * B * Nk * C) + 8 * B * N * M * K)
            d_w = d_w.view(B, N, M, K)
            return None, (d_w * attn.unsqueeze(3)).to(wdt), (d_w * nn_weight).sum(3).to(adt), d_val
        with torch.cuda.device(dev):
            _call("clusten_msdetrpc_bwd", dev, grad_feat.data_ptr(), nn_idx.data_ptr(), nn_weight.data_ptr(), attn.data_ptr(), val.data_ptr(),
                  off.data_ptr(), ent.data_ptr(), d_weight.data_ptr(), d_attn.data_ptr(), d_val.data_ptr(), B, N, Nk, C, M, K,
                  val.stride(0), val.stride(1), d_val.stride(0), d_val.stride(1), _lib.dtype_code(val),
                  nbytes=val.element_size() * (B * N * C + 2 * B * N * M * K + 2 * B * N * M + 2 * B * Nk * C) + 8 * B * N * M * K)
        return None, d_weight.to(wdt), d_attn.to(adt), d_val


# ---- row gather ------------------------------------------------------------------------------------------------------------
def gather_rows(src, idx):
    """``src.gather(1, idx.expand(-1, -1, src.shape[2]))`` for src [B, n, c] and idx int64 [B, k, 1] (the row reorders and
    selections of backbone/aff.py:332,335,340,471) through clusten_gather_rows: one index load per ROW and 16-byte copies instead
    of ATen's element-wise gather over the expanded index.  Any dtype (rows are opaque bytes).  No autograd: a source that needs a
    gradient, or a layout the kernel does not take, goes through ``torch.gather`` (same values)."""
    c = src.shape[2] if src.dim() == 3 else 0
    fast = (src.dim() == 3 and idx.dim() == 3 and idx.shape[2] == 1 and idx.shape[0] == src.shape[0] and src.is_cuda and idx.is_cuda
            and idx.dtype == torch.int64 and c > 0 and src.shape[1] > 0
            and not (torch.is_grad_enabled() and src.requires_grad)
            and src.shape[0] <= 0x7fffffff and src.shape[1] <= 0x7fffffff and idx.shape[1] <= 0x7fffffff
            and c * src.element_size() <= 0x7fffffff)
    if not fast:
        return src.gather(1, idx.expand(-1, -1, c) if src.dim() == 3 else idx)
    dev = _lib.require_cuda(src, idx)
    src, idx = src.contiguous(), idx.contiguous()
    B, n, k = src.shape[0], src.shape[1], idx.shape[1]
    out = torch.empty((B, k, c), dtype=src.dtype, device=dev)
    if B * k:
        rb = c * src.element_size()
        with torch.cuda.device(dev):
            _call("clusten_gather_rows", dev, src.data_ptr(), idx.data_ptr(), out.data_ptr(), B, n, k, rb, 0, nbytes=2 * B * k * rb + 8 * B * k)
    return out


# ---- stem: conv 3x3 / 2 + BatchNorm (eval) + GELU in one pass ------------------------------------------------------------------
STEM_OC = (16, 24, 32, 48, 64)


def stem_conv_bn_gelu_supported(x, conv, bn):
    return bool(x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and x.shape[1] == 3 and conv.in_channels == 3
                and conv.out_channels in STEM_OC and conv.kernel_size == (3, 3) and conv.stride == (2, 2) and conv.padding == (1, 1)
                and conv.dilation == (1, 1) and conv.groups == 1 and conv.padding_mode == "zeros" and conv.weight.dtype == torch.float32
                and not bn.training and bn.track_running_stats and bn.running_mean is not None and x.shape[0] <= 65535
                and not torch.is_autocast_enabled()
                and not (torch.is_grad_enabled() and (x.requires_grad or conv.weight.requires_grad or (bn.weight is not None and bn.weight.requires_grad))))


def stem_conv_bn_gelu(x, conv, bn, channels_last=False):
    """``gelu(bn(conv(x)))`` for the first stem convolution (PatchEmbed.forward, backbone/aff.py:549) in fp32 inference:
    clusten_stem_conv_bn_gelu, one pass over the [B, OC, H/2, W/2] map instead of cuDNN conv + ATen bias add + cuDNN BatchNorm + ATen
    GELU.  ``channels_last``: the result is pixel-major [B, H/2, W/2, OC] (what ``stem_im2col`` reads).  Caller checks
    ``stem_conv_bn_gelu_supported`` first.  No autograd."""
    dev = _lib.require_cuda(x, conv.weight, conv.bias, bn.running_mean, bn.running_var, bn.weight, bn.bias)
    x = x.contiguous()
    B, IC, H, W = x.shape
    OC = conv.out_channels
    OH, OW = (H + 1) // 2, (W + 1) // 2
    y = torch.empty((B, OH, OW, OC) if channels_last else (B, OC, OH, OW), dtype=torch.float32, device=dev)
    if y.numel():
        w = conv.weight.detach().contiguous()
        f = lambda t: None if t is None else t.detach().float().contiguous()  # noqa: E731
        cb, m, v, g, b = f(conv.bias), f(bn.running_mean), f(bn.running_var), f(bn.weight), f(bn.bias)
        with torch.cuda.device(dev):
            _call("clusten_stem_conv_bn_gelu", dev, x.data_ptr(), w.data_ptr(), _lib.ptr(cb), m.data_ptr(), v.data_ptr(), _lib.ptr(g),
                  _lib.ptr(b), float(bn.eps), y.data_ptr(), B, IC, H, W, OC, int(bool(channels_last)), nbytes=4 * (x.numel() + y.numel()))
    return y


def stem_im2col(mid, Kp):
    """Rows of a 3x3 / stride 2 / padding 1 convolution over the pixel-major map ``mid`` [B, H, W, C] (clusten_stem_im2col):
    A [B * OH * OW, Kp] with A[(b, py, px), (ky * 3 + kx) * C + c] = mid[b, 2 py - 1 + ky, 2 px - 1 + kx, c], zero outside the map and in
    the padding columns 9 C .. Kp - 1."""
    dev = _lib.require_cuda(mid)
    B, H, W, C = mid.shape
    OH, OW = (H + 1) // 2, (W + 1) // 2
    A = torch.empty((B * OH * OW, Kp), dtype=torch.float32, device=dev)
    if A.numel():
        with torch.cuda.device(dev):
            _call("clusten_stem_im2col", dev, mid.data_ptr(), A.data_ptr(), B, H, W, C, Kp, nbytes=4 * (mid.numel() + A.numel()))
    return A


_stem_w_cache = {}


def stem_proj2_weight(conv):
    """The second stem convolution's weight [E, C, 3, 3] as the GEMM operand of ``stem_im2col`` rows: [E, Kp] with column
    (ky * 3 + kx) * C + c, zero-padded to the 32-wide K chunk of the tcgen05 Linear; cached until the weight changes."""
    import weakref
    w = conv.weight
    hit = _stem_w_cache.get(id(w))
    if hit is not None and hit[0]() is w and hit[1] == w._version and hit[2] == w.data_ptr():
        return hit[3]
    E, C = w.shape[0], w.shape[1]
    Kp = (9 * C + 31) // 32 * 32
    w2 = torch.zeros((E, Kp), dtype=torch.float32, device=w.device)
    w2[:, :9 * C] = w.detach().permute(0, 2, 3, 1).reshape(E, 9 * C)
    if len(_stem_w_cache) > 256:
        _stem_w_cache.clear()
    _stem_w_cache[id(w)] = (weakref.ref(w), w._version, w.data_ptr(), w2)
    return w2


def stem_gemm_supported(x, conv1, bn, conv2):
    """The whole fp32 inference stem on our kernels: conv1 + BatchNorm + GELU (pixel-major), im2col, conv2 as a tcgen05 GEMM."""
    return bool(stem_conv_bn_gelu_supported(x, conv1, bn) and conv2.in_channels == conv1.out_channels and conv2.kernel_size == (3, 3)
                and conv2.stride == (2, 2) and conv2.padding == (1, 1) and conv2.dilation == (1, 1) and conv2.groups == 1
                and conv2.padding_mode == "zeros" and conv2.weight.dtype == torch.float32 and conv2.out_channels % 4 == 0
                and not (torch.is_grad_enabled() and conv2.weight.requires_grad))


def stem_tokens(x, conv1, bn, conv2):
    """``proj2(act1(bn(proj1(x)))).flatten(2).transpose(1, 2)`` (PatchEmbed.forward, backbone/aff.py:549-553) -> (tokens [B, h * w, E],
    h, w): conv1 + BatchNorm + GELU in one pass (pixel-major), the rows of the second convolution gathered once (``stem_im2col``),
    and that convolution as ONE GEMM on the tcgen05 Linear kernel (TF32 form of the split: its input is not range-bounded) whose
    output rows are the tokens -- no bias pass, no NCHW -> token-major copy.  Caller checks ``stem_gemm_supported``."""
    mid = stem_conv_bn_gelu(x, conv1, bn, channels_last=True)
    B, H2, W2, _ = mid.shape
    w2 = stem_proj2_weight(conv2)
    A = stem_im2col(mid, w2.shape[1])
    h, w = (H2 + 1) // 2, (W2 + 1) // 2
    y = linear_tc(A, w2, conv2.bias, "bias", split="tf32")
    return y.view(B, h * w, conv2.out_channels), h, w


# ---- rows of the relative-position feature table -------------------------------------------------------------------------------
def rel_pos_feature_rows(rows):
    """(dx, dy, dist, dy / dist, dx / dist) of the table rows ``rows`` (int64, any shape) -> fp32 [..., 5]: the reference's
    ``pre_table`` (backbone/aff.py:21-31) restricted to those rows, one kernel (clusten_rel_pos_features)."""
    dev = _lib.require_cuda(rows)
    r = rows.contiguous()
    out = torch.empty((*r.shape, 5), dtype=torch.float32, device=dev)
    if r.numel():
        with torch.cuda.device(dev):
            _call("clusten_rel_pos_features", dev, r.data_ptr(), out.data_ptr(), r.numel(), nbytes=28 * r.numel())
    return out
