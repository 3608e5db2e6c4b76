"""Convert between the reference's PyTables/HDF5 EMASE alignment file and the `.npz` twin.

    python -m gbrs_b200.convert in.h5 out.npz        # needs PyTables where it runs
    python -m gbrs_b200.convert in.npz out.h5        # writes the reference's layout (needs PyTables)

Both directions go through `gbrs_b200.AlignmentPropertyMatrix`, which reads whichever format the file is in (HDF5 files
need `tables` importable) and writes HDF5 only when PyTables is present and the name does not end in `.npz`.
"""
from __future__ import annotations

import sys

from .apm import AlignmentPropertyMatrix, _try_import_tables


def convert(src: str, dst: str, incidence_only: bool = True) -> None:
    apm = AlignmentPropertyMatrix(h5file=src)
    if not dst.endswith(".npz") and _try_import_tables() is None:
        raise RuntimeError("writing the HDF5 layout needs PyTables; give the output a .npz name instead")
    apm.save(dst, incidence_only=incidence_only)


def main(argv=None) -> int:
    argv = sys.argv[1:] if argv is None else argv
    if len(argv) != 2:
        print(__doc__)
        return 2
    convert(argv[0], argv[1])
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
