"""TEST INFRASTRUCTURE ONLY -- writes tests/golden/compress_*.npz by running the UNMODIFIED reference `compress`
(/root/reference/src/gbrs/gbrs/emase_utils.py:22-107) in the build container, through oracle/ref_harness.py's
pickle-backed stand-in for the missing PyTables.  Usage:  python -m oracle.make_golden_compress
"""
from __future__ import annotations

import os
import tempfile

import numpy as np
import scipy.sparse as sp

from oracle import ref_harness as rh

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def make_reads(seed, T, H, n_classes, n_reads, with_count, empty_reads=0):
    """Read-level incidence: `n_reads` reads drawn (with repetition) from `n_classes` random alignment patterns, plus
    some reads without any alignment.  Returns H CSC matrices (reads x loci) and a count vector or None."""
    rng = np.random.default_rng(seed)
    pats = []
    for _ in range(n_classes):
        k = int(min(1 + rng.poisson(1.2), 5))
        loci = np.unique((int(rng.integers(0, T)) + rng.integers(0, 4, k)) % T)
        pats.append([(int(t), int(rng.integers(1, 1 << H))) for t in loci])
    pick = rng.integers(0, n_classes, n_reads)
    rows = [[] for _ in range(H)]
    cols = [[] for _ in range(H)]
    order = list(pick) + [-1] * empty_reads
    rng.shuffle(order)
    for r, c in enumerate(order):
        if c < 0:
            continue
        for t, m in pats[c]:
            for h in range(H):
                if (m >> h) & 1:
                    rows[h].append(r)
                    cols[h].append(t)
    n = len(order)
    mats = [sp.csc_matrix((np.ones(len(rows[h])), (rows[h], cols[h])), shape=(n, T)) for h in range(H)]
    count = rng.integers(1, 6, n).astype(np.float64) if with_count else None
    return mats, count


def save_reference_file(ref, path, mats, count, hname, lname):
    T, H, n = mats[0].shape[1], len(mats), mats[0].shape[0]
    apm = ref.APM(shape=(T, H, n), haplotype_names=hname, locus_names=lname)
    for h in range(H):
        apm.data[h] = mats[h]
    apm.finalize()
    apm.count = count
    apm.save(h5file=path)


def run_case(name, specs, T, H):
    ref = rh.load_reference()
    hname = [chr(ord("A") + h) for h in range(H)]
    lname = [f"T{t:04d}" for t in range(T)]
    out = {"T": T, "H": H, "n_files": len(specs)}
    with tempfile.TemporaryDirectory() as tmp:
        paths = []
        for i, spec in enumerate(specs):
            mats, count = make_reads(T=T, H=H, **spec)
            p = os.path.join(tmp, f"in{i}.h5")
            save_reference_file(ref, p, mats, count, hname, lname)
            paths.append(p)
            for h in range(H):
                out[f"in{i}_h{h}_indptr"] = mats[h].indptr.astype(np.int64)
                out[f"in{i}_h{h}_indices"] = mats[h].indices.astype(np.int64)
            out[f"in{i}_n"] = mats[0].shape[0]
            out[f"in{i}_count"] = np.zeros(0) if count is None else count
            out[f"in{i}_has_count"] = count is not None
        outp = os.path.join(tmp, "out.h5")
        ref.gutils.compress(emase_files=paths, output_file=outp)  # the unmodified reference
        res = ref.APM(h5file=outp)
        out["n_ec"] = res.num_reads
        out["ec_count"] = np.asarray(res.count, dtype=np.float64)
        for h in range(H):
            m = sp.csc_matrix(res.data[h])
            m.sort_indices()
            out[f"ec_h{h}_indptr"] = m.indptr.astype(np.int64)
            out[f"ec_h{h}_indices"] = m.indices.astype(np.int64)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "reads", [out[f"in{i}_n"] for i in range(len(specs))], "->", out["n_ec"], "classes")


def main():
    os.makedirs(OUT, exist_ok=True)
    run_case("compress_one_file", [dict(seed=1, n_classes=60, n_reads=400, with_count=False, empty_reads=3)], T=40, H=8)
    run_case("compress_two_files_counts", [dict(seed=2, n_classes=80, n_reads=500, with_count=True),
                                           dict(seed=3, n_classes=80, n_reads=300, with_count=True, empty_reads=2)], T=50, H=3)
    run_case("compress_h1", [dict(seed=4, n_classes=30, n_reads=200, with_count=False)], T=25, H=1)


if __name__ == "__main__":
    main()
