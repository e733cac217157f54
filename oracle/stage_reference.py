"""Stage the UNMODIFIED reference under git-ignored `oracle/_ref/code/` so that it travels to the GPU box.
TEST / BASELINE INFRASTRUCTURE ONLY -- nothing under `sph-nca_b200/` reads it.

The reference is pure Python (Numba-CUDA kernels are JIT-compiled at run time): there is nothing to
compile, "building" `oracle/_ref` is a verbatim file copy from `/root/reference/code` made in the build
container by `__graft_entry__.build()`.  `oracle/_ref/` is listed in `.gitignore` (the sources never enter
this repository's history) and NOT in `.gpurunignore`, so the copy is present on the GPU box where
`/root/reference` does not exist.  Consumers:

* `oracle/ref_gpu_baseline.py` -- baseline #2 of BASELINE.json's north star: the reference's Numba-CUDA kernels
  + eager torch through its unchanged `nca.sample_plane`, timed on one B200 (started by `bench.py`);
* `tests/test_reference_drivers.py` -- runs the reference's own `losses.py` / `train.py` loop pieces on top of this
  repo's `sphops` + `nca` ("train.py runs unchanged" as evidence).

    python oracle/stage_reference.py        # idempotent; exit 0 and a note when /root/reference is absent
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("SPHNCA_REFERENCE_CODE", "/root/reference/code")
DST = os.path.join(HERE, "_ref", "code")
# what the two consumers import: the sphops package, the model / drivers, the losses, the geometry helpers, one texture
WANTED = ["sphops", "nca.py", "losses.py", "train.py", "test.py", "commons", os.path.join("data", "striped_0005.jpg")]


def stage(verbose=True):
    if not os.path.isdir(SRC):
        if verbose:
            print(f"stage_reference: {SRC} not present (GPU box?) -- keeping whatever is staged under {DST}")
        return os.path.isdir(DST)
    os.makedirs(DST, exist_ok=True)
    for rel in WANTED:
        s, d = os.path.join(SRC, rel), os.path.join(DST, rel)
        if not os.path.exists(s):
            continue
        if os.path.isdir(s):
            shutil.copytree(s, d, dirs_exist_ok=True, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
        else:
            os.makedirs(os.path.dirname(d), exist_ok=True)
            shutil.copy2(s, d)
    if verbose:
        print(f"stage_reference: staged {', '.join(WANTED)} -> {DST}")
    return True


if __name__ == "__main__":
    sys.exit(0 if stage() else 0)
