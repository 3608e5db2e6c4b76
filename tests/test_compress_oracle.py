"""The CPU restatement of `compress` (oracle/compress_oracle.py) against golden vectors produced by the unmodified
reference (oracle/make_golden_compress.py): identical classes in identical (first-appearance) order, identical counts."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

from oracle import compress_oracle as co

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
CASES = ["compress_one_file", "compress_two_files_counts", "compress_h1"]


def load_case(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    T, H = int(g["T"]), int(g["H"])
    files = []
    for i in range(int(g["n_files"])):
        n = int(g[f"in{i}_n"])
        mats = [sp.csc_matrix((np.ones(len(g[f"in{i}_h{h}_indices"])), g[f"in{i}_h{h}_indices"], g[f"in{i}_h{h}_indptr"]),
                              shape=(n, T)) for h in range(H)]
        files.append((mats, g[f"in{i}_count"] if bool(g[f"in{i}_has_count"]) else None))
    want = [sp.csc_matrix((np.ones(len(g[f"ec_h{h}_indices"])), g[f"ec_h{h}_indices"], g[f"ec_h{h}_indptr"]),
                          shape=(int(g["n_ec"]), T)) for h in range(H)]
    return T, H, files, want, g["ec_count"]


def same_pattern(a, b):
    a, b = sp.csc_matrix(a), sp.csc_matrix(b)
    a.sort_indices(); b.sort_indices()
    return a.shape == b.shape and np.array_equal(a.indptr, b.indptr) and np.array_equal(a.indices, b.indices)


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_compress(name):
    T, H, files, want, want_count = load_case(name)
    mats, count = co.compress(files)
    assert np.array_equal(count, want_count)  # integer-valued sums: exact
    for h in range(H):
        assert same_pattern(mats[h], want[h])
