"""Build recipe for the oracle's C restatement (gcc only).  TEST INFRASTRUCTURE ONLY.

The reference is pure Python (no native sources), so there is nothing to compile into
``oracle/_ref``; the reference's runnable Python is pinned through tests/golden/ instead
(oracle/gen_golden.py).  Output: oracle/_build/libcindex_oracle.so (git-ignored, travels
to the GPU box with the snapshot).
"""
from __future__ import annotations

import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "cindex_oracle.c")
    out_dir = os.path.join(_HERE, "_build")
    out = os.path.join(out_dir, "libcindex_oracle.so")
    os.makedirs(out_dir, exist_ok=True)
    if (not force and os.path.exists(out) and os.path.getmtime(out) >= os.path.getmtime(src)):
        return out
    cmd = ["gcc", "-O2", "-fopenmp", "-fno-fast-math", "-shared", "-fPIC", "-o", out, src, "-lm"]
    subprocess.check_call(cmd)
    return out


if __name__ == "__main__":
    print(build(force=True))
