"""The native table writer must print every double exactly like python's repr / str(numpy.float64)."""
import io
import os

import numpy as np
import pytest

from gbrs_b200 import utils


def test_float_formatting_matches_python_repr():
    rng = np.random.default_rng(3)
    vals = np.concatenate([
        rng.random(60000) * 10.0 ** rng.integers(-30, 30, 60000),
        -rng.random(5000) * 10.0 ** rng.integers(-8, 8, 5000),
        np.round(rng.random(5000) * 1e6), np.round(rng.random(2000) * 50) / 8.0,
        np.array([0.0, -0.0, 1.0, 5.0, 1e16, 9999999999999998.0, 1e15, 1e-4, 9.999e-5, 1e-5, 1e-7, 123456789012345678.0,
                  2.5e-320, 5e-324, 1.7976931348623157e308, 1e22, 1e21, 0.1 + 0.2, 100.0, 1e17, 3.0e-5, 0.5, 1 / 3,
                  2 / 3, 1e100, 1.5e-100, 4.35, 0.3, 1000000.0, 123456.789, float("inf"), float("-inf")]),
    ])
    for v in vals:
        assert utils.py_float_str(v) == repr(float(v)), v
    assert utils.py_float_str(float("nan")) == "nan"


def test_write_table_rows_equals_python_loop(tmp_path):
    rng = np.random.default_rng(4)
    n = 5000
    names = [f"T{t:05d}" for t in range(n)]
    data = rng.random((9, n)) * 10.0 ** rng.integers(-6, 9, (9, n))
    data[:, ::7] = 0.0
    data[3, :] = np.round(data[3, :])
    notes = {nm: ("AB" if i % 3 else None) for i, nm in enumerate(names)}
    order = np.argsort(data[-1])[::-1]
    for use_notes, use_order in [(False, False), (True, False), (True, True)]:
        path = os.path.join(str(tmp_path), f"t{int(use_notes)}{int(use_order)}.tsv")
        with open(path, "w") as fh:
            fh.write("locus\tA\ttotal\n")
            utils.write_table_rows(fh, names, data, notes=notes if use_notes else None,
                                   order=order if use_order else None)
            fh.write("tail\n")  # the python file object keeps working after the native append
        exp = io.StringIO()
        exp.write("locus\tA\ttotal\n")
        for i in (order if use_order else range(n)):
            line = names[i] + "\t" + "\t".join(str(np.float64(data[k, i])) for k in range(9))
            if use_notes:
                line += f"\t{notes[names[i]]}"
            exp.write(line + "\n")
        exp.write("tail\n")
        assert open(path).read() == exp.getvalue()
