"""Stage the few reference sources the GPU-side drop-in test executes (TEST INFRASTRUCTURE).

``tests/test_gpu_dropin.py`` runs the reference's OWN ``AFF`` class on the B200 with ``..clusten`` bound to this repository's ops.
/root/reference does not exist on the GPU box, so the files the test executes are copied, unmodified, into the git-ignored
``baseline/_ref/`` (it travels with the gpurun snapshot, it is never committed and nothing in the product package imports it):

    mask2former/modeling/backbone/aff.py            the AFF backbone (caller of CLUSTENQK / AV / WF, aff.py:114,154,361)
    mask2former/modeling/backbone/point_utils.py    space_filling_cluster, knn_keops call sites, upsample_feature_shepard

Usage:  python oracle/stage_ref.py        (run by __graft_entry__.build() when /root/reference is present)
"""
import os
import shutil

SRC = "/root/reference/mask2former/modeling"
DST = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref", "mask2former", "modeling")
FILES = ["backbone/aff.py", "backbone/point_utils.py"]


def stage():
    if not os.path.isdir(SRC):
        print(f"[stage_ref] {SRC} absent (GPU box?) -- using the staged copy only")
        return False
    for rel in FILES:
        dst = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(SRC, rel), dst)
    print(f"[stage_ref] staged {len(FILES)} reference files under {DST}")
    return True


if __name__ == "__main__":
    stage()
