"""TEST INFRASTRUCTURE ONLY -- writes tests/golden/stencil_small.npz by running the UNMODIFIED reference `stencil`
(/root/reference/src/gbrs/gbrs/emase_utils.py:110-177) in the build container through oracle/ref_harness.py's
stand-in for PyTables.  Usage:  python -m oracle.make_golden_stencil
"""
from __future__ import annotations

import os
import tempfile

import numpy as np
import scipy.sparse as sp

from gbrs_b200 import synth
from oracle import ref_harness as rh

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def main():
    ref = rh.load_reference()
    d = synth.generate(T=60, N=700, H=8, with_genotype=True, sample_index=31)
    out = {"T": d.T, "H": d.H, "N": d.N}
    with tempfile.TemporaryDirectory() as tmp:
        aln, grp, gt, res = (os.path.join(tmp, x) for x in ("aln.h5", "grp.tsv", "gt.tsv", "out.h5"))
        apm = rh.build_reference_apm(d)
        apm.save(h5file=aln)
        synth.write_group_file(d, grp)
        # drop two genes from the genotype file (they are masked out completely) and add a comment line
        with open(gt, "w") as fh:
            fh.write("#Gene_ID\tDiplotype\n")
            for g, call in enumerate(d.genotype):
                if g not in (3, 11):
                    fh.write(f"{d.gname[g]}\t{call}\n")
        out["genotype_tsv"] = np.array(open(gt).read())
        out["group_tsv"] = np.array(open(grp).read())
        ref.gutils.stencil(alignment_file=aln, genotype_file=gt, group_file=grp, output_file=res)  # unmodified reference
        got = ref.APM(h5file=res)
        out["count"] = np.asarray(got.count, dtype=np.float64)
        for h in range(d.H):
            m = sp.csc_matrix(got.data[h])
            m.sort_indices()
            out[f"out_h{h}_indptr"] = m.indptr.astype(np.int64)
            out[f"out_h{h}_indices"] = m.indices.astype(np.int64)
    out["pair_class"], out["pair_locus"], out["pair_mask"], out["in_count"] = d.pair_class, d.pair_locus, d.pair_mask, d.count
    out["sample_index"] = 31
    np.savez_compressed(os.path.join(OUT, "stencil_small.npz"), **out)
    print("stencil_small: nnz", d.nnz, "->", sum(len(out[f"out_h{h}_indices"]) for h in range(d.H)))


if __name__ == "__main__":
    main()
