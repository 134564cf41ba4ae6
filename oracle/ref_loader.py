"""Import the reference's OWN Python (point_utils.py, aff.py) on CPU -- authoring container only.

TEST INFRASTRUCTURE.  /root/reference does not exist on the GPU box, so nothing at test / smoke /
bench run time imports this module; it is used by oracle/make_golden.py to generate the committed
fixtures in tests/golden and by the (skipped-when-absent) oracle-vs-reference tests.

Recipe: the reference files are loaded by path under a fake package
``_affref`` so their relative imports resolve; ``detectron2``, ``timm`` and ``pykeops`` (not
installed, no network) are replaced by minimal stubs; ``..clusten`` is bound to the CPU
restatements of oracle/clusten_ops.py; ``knn_keops`` is replaced (pykeops cannot be imported)
by oracle.point_ops.knn; torch.sort / topk inside the reference are forced to the canonical
stable forms while ``canonical_ties()`` is active.
"""
import contextlib
import importlib.util
import os
import sys
import types

import torch

REF_ROOT = "/root/reference/mask2former/modeling"
# git-ignored copy of the few reference files the drop-in test executes, staged by oracle/stage_ref.py so that it travels to the
# GPU box (never committed, never imported by the product package)
STAGED_ROOT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref", "mask2former", "modeling")


def available():
    return os.path.isfile(os.path.join(REF_ROOT, "backbone", "aff.py"))


def dropin_root():
    """Where the reference's backbone sources can be executed from: the reference tree, else the staged copy, else None."""
    for root in (REF_ROOT, STAGED_ROOT):
        if os.path.isfile(os.path.join(root, "backbone", "aff.py")):
            return root
    return None


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _install_stubs():
    class _Registry:
        def register(self, obj=None):
            return obj if obj is not None else (lambda o: o)

    class DropPath(torch.nn.Module):            # timm 0.6.12 DropPath: identity in eval / p=0
        def __init__(self, p=0.0):
            super().__init__()
            self.p = p

        def forward(self, x):
            return x

    class ShapeSpec:
        def __init__(self, channels=None, stride=None, **kw):
            self.channels, self.stride = channels, stride

    if "timm" not in sys.modules:
        _mod("timm"), _mod("timm.models")
        _mod("timm.models.layers", DropPath=DropPath, trunc_normal_=torch.nn.init.trunc_normal_)
    if "detectron2" not in sys.modules:
        _mod("detectron2")
        _mod("detectron2.modeling", BACKBONE_REGISTRY=_Registry(), SEM_SEG_HEADS_REGISTRY=_Registry(),
             Backbone=torch.nn.Module, ShapeSpec=ShapeSpec)


_cache = {}


def load():
    """Returns (point_utils, aff) reference modules wired to the CPU oracle ops."""
    if "mods" in _cache:
        return _cache["mods"]
    if not available():
        raise RuntimeError("reference tree not present (GPU box) -- use tests/golden fixtures")
    _install_stubs()
    from . import clusten_ops, point_ops

    pkg = _mod("_affref")
    pkg.__path__ = []
    cl = _mod("_affref.clusten",
              CLUSTENQKFunction=clusten_ops.CLUSTENQKFunction,
              CLUSTENAVFunction=clusten_ops.CLUSTENAVFunction,
              CLUSTENWFFunction=clusten_ops.CLUSTENWFFunction,
              WEIGHTEDGATHERFunction=clusten_ops.WEIGHTEDGATHERFunction,
              MSDETRPCFunction=clusten_ops.MSDETRPCFunction)
    pkg.clusten = cl
    bb = _mod("_affref.backbone")
    bb.__path__ = []

    def _load(name, rel):
        spec = importlib.util.spec_from_file_location(name, os.path.join(REF_ROOT, rel))
        m = importlib.util.module_from_spec(spec)
        sys.modules[name] = m
        spec.loader.exec_module(m)
        return m

    pu = _load("_affref.backbone.point_utils", "backbone/point_utils.py")
    aff = _load("_affref.backbone.aff", "backbone/aff.py")
    # pykeops is not importable: replace knn_keops in every module that imported it by name
    pu.knn_keops = point_ops.knn
    aff.knn_keops = point_ops.knn
    _cache["mods"] = (pu, aff)
    return pu, aff


def load_dropin():
    """The reference's OWN ``aff.py`` / ``point_utils.py`` with ``..clusten`` bound to THIS REPOSITORY'S CUDA ops and ``knn_keops`` to
    its kNN -- the literal drop-in of clusten/__init__.py:6 and the call sites aff.py:114,154,361: returns (point_utils, aff).
    The modules live under their own fake package, apart from the CPU-oracle binding of ``load()``."""
    if "dropin" in _cache:
        return _cache["dropin"]
    root = dropin_root()
    if root is None:
        raise RuntimeError("no reference backbone sources (neither /root/reference nor baseline/_ref; run oracle/stage_ref.py)")
    _install_stubs()
    import autofocusformermod_b200 as P

    pkg = _mod("_affdrop")
    pkg.__path__ = []
    pkg.clusten = _mod("_affdrop.clusten", CLUSTENQKFunction=P.CLUSTENQKFunction, CLUSTENAVFunction=P.CLUSTENAVFunction,
                       CLUSTENWFFunction=P.CLUSTENWFFunction, WEIGHTEDGATHERFunction=P.WEIGHTEDGATHERFunction,
                       MSDETRPCFunction=P.MSDETRPCFunction)
    bb = _mod("_affdrop.backbone")
    bb.__path__ = []

    def _load(name, rel):
        spec = importlib.util.spec_from_file_location(name, os.path.join(root, rel))
        m = importlib.util.module_from_spec(spec)
        sys.modules[name] = m
        spec.loader.exec_module(m)
        return m

    pu = _load("_affdrop.backbone.point_utils", "backbone/point_utils.py")
    aff = _load("_affdrop.backbone.aff", "backbone/aff.py")
    pu.knn_keops = P.knn_keops                      # pykeops is not importable: the repo's kNN under the reference's name
    aff.knn_keops = P.knn_keops
    _cache["dropin"] = (pu, aff)
    return pu, aff


def load_point_conv():
    """The reference's own ``PointConv`` class (pixel_decoder/msdeformattn_pc.py:271-314), executed from the reference file
    without importing the rest of the module (which needs detectron2 / fvcore): the class source is cut out with ``ast``
    and run against the reference ``aff`` module's table constants, the oracle kNN and the oracle CLUSTENWF."""
    import ast
    pu, aff = load()
    from . import clusten_ops, point_ops
    path = os.path.join(REF_ROOT, "pixel_decoder", "msdeformattn_pc.py")
    src = open(path).read()
    node = next(n for n in ast.parse(src).body if isinstance(n, ast.ClassDef) and n.name == "PointConv")
    ns = {"nn": torch.nn, "torch": torch, "knn_keops": point_ops.knn, "CLUSTENWFFunction": clusten_ops.CLUSTENWFFunction,
          "pre_table": aff.pre_table, "rel_pos_width": aff.rel_pos_width, "table_width": aff.table_width}
    exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), ns)
    return ns["PointConv"]


def load_msdeformattn_pc():
    """The reference's own ``MSDeformAttnPc`` class and ``scale_pos`` (pixel_decoder/msdeformattn_pc.py:28-53, 107-205), cut out with
    ``ast`` and run against the reference ``point_utils`` (oracle kNN) and the oracle MSDETRPC restatement."""
    import ast
    import math
    pu, _ = load()
    from . import clusten_ops
    path = os.path.join(REF_ROOT, "pixel_decoder", "msdeformattn_pc.py")
    body = ast.parse(open(path).read()).body
    nodes = [n for n in body if (isinstance(n, ast.ClassDef) and n.name == "MSDeformAttnPc") or (isinstance(n, ast.FunctionDef) and n.name == "scale_pos")]
    ns = {"nn": torch.nn, "torch": torch, "F": torch.nn.functional, "math": math, "constant_": torch.nn.init.constant_,
          "xavier_uniform_": torch.nn.init.xavier_uniform_, "upsample_feature_shepard": pu.upsample_feature_shepard,
          "MSDETRPCFunction": clusten_ops.MSDETRPCFunction}
    exec(compile(ast.Module(body=nodes, type_ignores=[]), path, "exec"), ns)
    return ns["MSDeformAttnPc"], ns["scale_pos"]


def load_point2img():
    """The reference's own ``point2img`` (transformer_decoder/mask2former_transformer_decoder.py:20-39), cut out with ``ast``."""
    import ast
    path = os.path.join(REF_ROOT, "transformer_decoder", "mask2former_transformer_decoder.py")
    node = next(n for n in ast.parse(open(path).read()).body if isinstance(n, ast.FunctionDef) and n.name == "point2img")
    ns = {"torch": torch}
    exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), ns)
    return ns["point2img"]


@contextlib.contextmanager
def canonical_ties():
    """Force the canonical tie rules inside reference code: Tensor.sort -> stable,
    Tensor.topk(sorted=False) -> first k of a stable descending sort."""
    orig_sort, orig_topk = torch.Tensor.sort, torch.Tensor.topk

    def sort(self, *a, **kw):
        kw["stable"] = True
        if a:
            kw["dim"] = a[0]
            if len(a) > 1:
                kw["descending"] = a[1]
        return orig_sort(self, **kw)

    def topk(self, k, dim=-1, largest=True, sorted=True):
        v, i = orig_sort(self, dim=dim, descending=largest, stable=True)
        return v.narrow(dim, 0, k), i.narrow(dim, 0, k)

    torch.Tensor.sort, torch.Tensor.topk = sort, topk
    try:
        yield
    finally:
        torch.Tensor.sort, torch.Tensor.topk = orig_sort, orig_topk
