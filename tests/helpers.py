"""Shared helpers for the test-suite (CPU side only: golden loading, oracle wrappers)."""
from __future__ import annotations

import glob
import os

import numpy as np

from gbrs_b200 import synth
from oracle import em_oracle as eo

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_em_cases():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "em_*.npz")))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    g = {k: z[k] for k in z.files}
    for k in ("T", "N", "H", "model", "iters", "max_iters", "n_genes"):
        g[k] = int(g[k])
    for k in ("pseudocount", "tol"):
        g[k] = float(g[k])
    g["masked"] = bool(g["masked"])
    return g


def synth_from_golden(g):
    """SynthData for a golden case: stored inputs if present, else regenerated from the seed."""
    if "pair_class" in g:
        d = synth.SynthData(T=g["T"], H=g["H"], N=g["N"], pair_class=g["pair_class"].astype(np.int64),
                            pair_locus=g["pair_locus"].astype(np.int64), pair_mask=g["pair_mask"],
                            count=g["count"], gene_of=g["gene_of"].astype(np.int64), lengths=g["lengths"],
                            hname=synth.HAPLOTYPES[:g["H"]])
        d.lname = [f"T{t:07d}" for t in range(d.T)]
        d.gname = [f"G{x:07d}" for x in range(int(d.gene_of.max()) + 1)]
        return d
    return synth.generate(T=g["T"], N=g["N"], H=g["H"], with_genotype=g["masked"],
                          n_genes=None if g["n_genes"] < 0 else g["n_genes"])


def oracle_run(d, model, pseudocount=0.0, tol=1e-4, max_iters=999, gtmask=None):
    apm = eo.apm_from_pairs(d.T, d.H, d.N, d.pair_class, d.pair_locus, d.pair_mask, d.count)
    if gtmask is not None:
        apm = eo.apply_genotype_mask(apm, gtmask)
    eff = eo.effective_length_table(d.lengths)
    gene_of = eo.gene_index(d.T, d.groups())
    theta0 = eo.prepare(apm, eff, pseudocount)
    out = eo.run(apm, theta0, model, eff, gene_of, tol=tol, max_iters=max_iters)
    out["theta0"] = theta0
    return out


def relerr(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    scale = np.abs(b).max()
    return float(np.abs(a - b).max() / scale) if scale > 0 else float(np.abs(a).max())


def elementwise_relerr(a, b, floor):
    """max |a-b| / max(|b|, floor): the 1e-6 bar is applied element-wise with an absolute floor so that
    entries that are (near) zero in both do not divide by zero."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float((np.abs(a - b) / np.maximum(np.abs(b), floor)).max())
