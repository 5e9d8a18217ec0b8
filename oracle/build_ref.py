#!/usr/bin/env python
"""Recipe for `oracle/_ref/`: the UNMODIFIED reference's Python packages, staged so that they travel to the GPU box.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  The reference is pure Python with no setup.py / pyproject
(SURVEY.md 0), so "building" it is staging the four packages the learner path imports -- `agents`, `networks`, `utils`,
`buffers` -- from where they lie under /root/reference into `oracle/_ref/` (git-ignored: no reference source enters the
history; NOT gpurun-ignored: the directory ships with the snapshot like a built .so).  `oracle/ref_driver.py` then runs
the reference's verbatim `learning()` coroutine from there: the `cpu_baseline` / `--impl reference` legs of bench.py
become `kind: "reference"`, and the in-run parity block compares the CUDA path with the reference itself.

    python oracle/build_ref.py          # no-op (returns False) when /root/reference is absent, e.g. on the GPU box
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
OUT = os.path.join(HERE, "_ref")
PACKAGES = ("agents", "networks", "utils", "buffers")


def build(force=False):
    if not os.path.isdir(REF):
        return os.path.isdir(OUT)
    stamp = os.path.join(OUT, ".staged")
    if os.path.exists(stamp) and not force:
        return True
    if os.path.isdir(OUT):
        shutil.rmtree(OUT)
    os.makedirs(OUT)
    for p in PACKAGES:
        shutil.copytree(os.path.join(REF, p), os.path.join(OUT, p),
                        ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    with open(stamp, "w") as f:
        f.write("staged from /root/reference by oracle/build_ref.py (unmodified)\n")
    return True


if __name__ == "__main__":
    print("oracle/_ref:", "ready" if build(force="--force" in sys.argv) else "reference not available")
