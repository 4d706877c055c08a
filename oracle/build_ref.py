"""TEST / BENCH INFRASTRUCTURE — recipe that makes the UNMODIFIED reference available as ``oracle/_ref/``.

The reference (henrycharlesworth/settlers_of_catan_RL) is pure Python: there is nothing to compile.  This
recipe copies the three packages the hot path and its callers live in (``game/``, ``env/``, ``RL/`` incl. the
shipped checkpoint ``RL/results/default_after_update_3825.pt``) from ``/root/reference`` — where they lie,
unchanged — into the git-ignored ``oracle/_ref/``, so that they travel to the GPU box with the working tree
like a built ``.so`` does (``oracle/_ref/`` is in ``.gitignore``, not in ``.gpurunignore``).  Nothing of it
enters the history, and nothing of the product package may import it (``tests/test_abi.py`` checks that):

  * ``tests/`` pin the oracle against it (``tests/test_oracle_vs_reference.py`` etc.), on the box as well;
  * ``bench.py --impl reference`` and the ``cpu_baseline`` leg time its ``EnvWrapper`` loop on the box's
    host cores (``oracle/ref_bench.py``).

``__graft_entry__.build()`` runs this wherever ``/root/reference`` exists (this container); on the GPU box the
already-built copy is used.
"""
from __future__ import annotations

import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference"
DST = os.path.join(HERE, "_ref")
PACKAGES = ("game", "env", "RL")


def _newest(root: str) -> float:
    t = 0.0
    for d, _, files in os.walk(root):
        for f in files:
            t = max(t, os.path.getmtime(os.path.join(d, f)))
    return t


def build(force: bool = False) -> str | None:
    """copy the reference's packages into oracle/_ref (idempotent); returns the path, or None when neither the
    reference nor an earlier copy exists"""
    have_src = os.path.isdir(os.path.join(SRC, "game"))
    have_dst = os.path.isdir(os.path.join(DST, "game"))
    if not have_src:
        return DST if have_dst else None
    if have_dst and not force and all(
            os.path.isdir(os.path.join(DST, p)) and _newest(os.path.join(DST, p)) >= _newest(os.path.join(SRC, p)) for p in PACKAGES):
        return DST
    os.makedirs(DST, exist_ok=True)
    for p in PACKAGES:
        dst = os.path.join(DST, p)
        if os.path.isdir(dst):
            shutil.rmtree(dst)
        shutil.copytree(os.path.join(SRC, p), dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    with open(os.path.join(DST, "ORIGIN.txt"), "w") as f:
        f.write("verbatim copy of %s/{%s} made by oracle/build_ref.py; git-ignored; test and bench infrastructure only\n" % (SRC, ",".join(PACKAGES)))
    return DST


def root(copy_only: bool = False) -> str | None:
    """where the reference can be imported from: the mounted tree if present (tests in this container), else the
    copy; ``copy_only`` (bench.py, which must never read /root/reference): the copy or nothing"""
    if not copy_only and os.path.isdir(os.path.join(SRC, "game")):
        return SRC
    if os.path.isdir(os.path.join(DST, "game")):
        return DST
    return None


if __name__ == "__main__":
    print(build(force=True))
