"""ORACLE-side build step (test infrastructure): make the unmodified reference importable where /root/reference does
not exist (the GPU box).  The reference is pure Python, so "building" it is a verbatim copy of its `fs2` package into
`oracle/_ref/fs2` — git-ignored, never part of the repo history, shipped with the snapshot like a built `.so`.
`bench.py --impl reference` / `cpu_baseline` then time the REAL reference (dropout RNG included) on the host cores.
"""
from __future__ import annotations

import shutil
from pathlib import Path

SRC = Path("/root/reference/fs2")
DST = Path(__file__).resolve().parent / "_ref" / "fs2"


def build_ref() -> bool:
    """True when oracle/_ref/fs2 is in place (copied now or earlier)."""
    if not SRC.exists():
        return (DST / "model.py").exists()
    if DST.exists():
        shutil.rmtree(DST)
    shutil.copytree(SRC, DST, ignore=shutil.ignore_patterns("tests", "__pycache__", "*.pyc"))
    return True


if __name__ == "__main__":
    print(build_ref())
