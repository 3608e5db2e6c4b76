"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's `compress` (equivalence-class construction and
replicate merging), /root/reference/src/gbrs/gbrs/emase_utils.py:22-107.

Nothing under gbrs_b200/ imports this file; only tests/ (and oracle/make_golden_compress.py) do.  It is pinned against
the unmodified reference executed in the build container (tests/golden/compress_*.npz, oracle/make_golden_compress.py).

The reference walks every read of every input file in order, builds a string key from the sorted locus ids the read
hits in each haplotype (emase_utils.py:62-71: `':'.join(','.join(sorted loci of h) for h)`), sums the read counts per
key in a dict (:72) -- so classes are numbered in order of FIRST APPEARANCE -- and writes an incidence matrix with one
row per key (:92-100).  A read without any alignment has the key ':' * (H - 1) and forms a class with an empty row.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp


def read_keys(mats):
    """mats: H sparse matrices (reads x loci).  Returns per read a tuple of H tuples of sorted locus ids -- the same
    information as the reference's string key (emase_utils.py:62-71)."""
    csr = [sp.csr_matrix(m) for m in mats]
    for m in csr:
        m.sort_indices()
    n = csr[0].shape[0]
    keys = []
    for r in range(n):
        keys.append(tuple(tuple(int(x) for x in m.indices[m.indptr[r]:m.indptr[r + 1]]) for m in csr))
    return keys


def compress(files):
    """files: list of (mats, count_or_None) in input order.  Returns (ec_mats, ec_count): H CSC matrices (classes x
    loci, values 1.0) and the class counts, classes in order of first appearance (emase_utils.py:46-100)."""
    ec = {}
    shape = None
    for mats, count in files:
        n, T = mats[0].shape
        shape = (len(mats), T)
        if count is None:
            count = np.ones(n)  # emase_utils.py:58-59
        for r, key in enumerate(read_keys(mats)):
            ec[key] = ec.get(key, 0.0) + float(count[r])  # :72
    H, T = shape
    n_ec = len(ec)
    rows = [[] for _ in range(H)]
    cols = [[] for _ in range(H)]
    counts = np.zeros(n_ec)
    for row_id, (key, c) in enumerate(ec.items()):  # dict order = first appearance (:92)
        counts[row_id] = c
        for h in range(H):
            for t in key[h]:
                rows[h].append(row_id)
                cols[h].append(t)
    mats = [sp.csc_matrix((np.ones(len(rows[h])), (rows[h], cols[h])), shape=(n_ec, T)) for h in range(H)]
    return mats, counts
