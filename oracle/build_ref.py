"""Build recipe for ``oracle/_ref`` -- TEST INFRASTRUCTURE, never imported by the product package.

Compiles the reference's OWN CLUSTEN CUDA extensions for sm_100a straight from the sources where
they lie under ``/root/reference`` (nothing is copied into this repository) and leaves only the
built pybind modules in ``oracle/_ref/`` (git-ignored, still shipped to the GPU box by gpurun).

    reference sources                                              module
    mask2former/modeling/clusten/src/clustenqk_cuda{.cpp,_kernel.cu}        clustenqk_cuda
    mask2former/modeling/clusten/src/clustenav_cuda{.cpp,_kernel.cu}        clustenav_cuda
    mask2former/modeling/clusten/src/clustenwf_cuda{.cpp,_kernel.cu}        clustenwf_cuda
    mask2former/modeling/clusten/src/weighted_gather_cuda{.cpp,_kernel.cu}  weighted_gather_cuda

The reference's own build system (``clusten/src/setup.py``) is NOT run; this is our own short
recipe on top of ``torch.utils.cpp_extension.load`` (ninja + nvcc).  The modules cannot execute
in the authoring container (no GPU); on the GPU box ``tests/test_gpu_vs_reference_kernels.py``
loads them through ``oracle/ref_cuda.py`` as a second, independent oracle: the real reference
kernels run on the same seeded inputs as ours.

Usage:  python oracle/build_ref.py [module ...]        (default: all four)
"""
import os
import sys

REF_SRC = "/root/reference/mask2former/modeling/clusten/src"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
MODULES = ["clustenqk_cuda", "clustenav_cuda", "clustenwf_cuda", "weighted_gather_cuda"]


def build(mod):
    if not os.path.isdir(REF_SRC):
        print(f"[build_ref] {REF_SRC} absent (GPU box?) -- using prebuilt files only")
        return False
    os.environ["TORCH_CUDA_ARCH_LIST"] = "10.0a"
    os.environ.setdefault("MAX_JOBS", "2")
    from torch.utils.cpp_extension import load
    bdir = os.path.join(OUT, mod)
    os.makedirs(bdir, exist_ok=True)
    load(name=mod,
         sources=[os.path.join(REF_SRC, mod + ".cpp"), os.path.join(REF_SRC, mod + "_kernel.cu")],
         build_directory=bdir, verbose=False, is_python_module=False,
         extra_cuda_cflags=["-O3"])
    so = os.path.join(bdir, mod + ".so")
    print(f"[build_ref] {mod}: {'ok' if os.path.exists(so) else 'MISSING'} -> {so}")
    return os.path.exists(so)


if __name__ == "__main__":
    mods = sys.argv[1:] or MODULES
    ok = all([build(m) for m in mods])
    sys.exit(0 if ok else 1)
