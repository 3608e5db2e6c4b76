"""TEST / BASELINE INFRASTRUCTURE ONLY -- stages the reference's own implementation of the hot path into the git-ignored
directory oracle/_ref/ so that it travels to the GPU box (where /root/reference does not exist) and `bench.py --impl
reference` can time the UNMODIFIED reference there (`cpu_baseline.kind = "reference"`).

    python -m oracle.make_ref        (also run by __graft_entry__.build() whenever /root/reference is mounted)

What is staged, byte for byte, from /root/reference/src/gbrs: the three modules of the path (emase/EMfactory.py,
emase/AlignmentPropertyMatrix.py, emase/Sparse3DMatrix.py), the package `__init__` files and utils.py (imported by all
three).  Nothing under oracle/_ref/ is tracked by git, nothing in gbrs_b200/ imports it, and only bench.py's reference arm
and oracle/ref_harness.py read it."""
from __future__ import annotations

import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/src/gbrs"
DST = os.path.join(HERE, "_ref", "src", "gbrs")
FILES = ["__init__.py", "utils.py", "emase/__init__.py", "emase/EMfactory.py", "emase/AlignmentPropertyMatrix.py",
         "emase/Sparse3DMatrix.py"]


def stage(verbose: bool = True) -> bool:
    """Copy the files; False (and nothing touched) where the reference is not mounted."""
    if not os.path.isdir(SRC):
        return False
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(SRC, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest[rel] = hashlib.sha256(open(dst, "rb").read()).hexdigest()
    with open(os.path.join(HERE, "_ref", "MANIFEST.json"), "w") as fh:
        json.dump({"source": SRC, "sha256": manifest}, fh, indent=1)
    if verbose:
        print(f"staged {len(FILES)} reference files into {os.path.dirname(os.path.dirname(DST))}")
    return True


if __name__ == "__main__":
    if not stage():
        raise SystemExit(f"{SRC} is not present: nothing staged")
