"""TEST INFRASTRUCTURE ONLY -- writes tests/golden/reconstruct_*.npz by running the UNMODIFIED reference `reconstruct`
(/root/reference/src/gbrs/gbrs/gbrs_utils.py:382-609) in the build container.  The reference module imports matplotlib at
the top (not installed here, not used by `reconstruct`): an empty stand-in module is registered before the import.
Usage:  python -m oracle.make_golden_reconstruct

Each fixture holds the logical inputs (so the GPU box can rebuild the files without /root/reference) and the reference's
three outputs: the posterior per chromosome (`*.genoprobs.npz`), the ordered Viterbi states (`*.genotypes.npz`) and the
gene -> diplotype table (`*.genotypes.tsv`).
"""
from __future__ import annotations

import importlib
import os
import sys
import tempfile
import types

import numpy as np

from gbrs_b200 import synth

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
REFERENCE_SRC = "/root/reference/src"

CASES = {
    # name: (generator kwargs, reconstruct kwargs)
    "reconstruct_h8": (dict(genes_per_chrom=(22, 13, 1, 16), H=8, sample_index=0, extra_tprob_step=("2",)),
                       dict(expr_threshold=1.5, sigma=0.12)),
    "reconstruct_h2": (dict(genes_per_chrom=(60, 18), H=2, sample_index=3), dict(expr_threshold=1.0, sigma=0.2)),
    "reconstruct_h4": (dict(genes_per_chrom=(30,), H=4, sample_index=5, frac_low=0.4),
                       dict(expr_threshold=2.0, sigma=0.12)),
}


def load_reference_gbrs_utils(data_dir: str):
    """The reference reads $GBRS_DATA once, at import (gbrs_utils.py:18): set it, then (re)import."""
    os.environ["GBRS_DATA"] = data_dir
    for name in ("matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].use = lambda *a, **k: None
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)
    mod = importlib.import_module("gbrs.gbrs.gbrs_utils")
    return importlib.reload(mod)


def pack_inputs(d: synth.SynthReconstruct) -> dict:
    out = {"hname": np.array(d.hname), "chroms": np.array(d.chroms),
           "chrlen": np.array([d.chrlen[c] for c in d.chroms], dtype=np.int64),
           "expr_genes": np.array(list(d.expr.keys())), "expr": np.array(list(d.expr.values())),
           "avec_genes": np.array(list(d.avecs.keys())), "avecs": np.array(list(d.avecs.values()))}
    for c in d.genes:
        out[f"genes_{c}"] = np.array(d.genes[c])
        out[f"gpos_{c}"] = np.array(d.gpos[c], dtype=np.int64)
        out[f"tprob_{c}"] = d.tprob[c]
    return out


def unpack_inputs(z) -> synth.SynthReconstruct:
    """Inverse of pack_inputs (used by the tests; needs only the fixture)."""
    chroms = [str(c) for c in z["chroms"]]
    gchroms = [k[len("genes_"):] for k in z.files if k.startswith("genes_")]
    return synth.SynthReconstruct(
        hname=tuple(str(h) for h in z["hname"]), chroms=chroms,
        chrlen={c: int(n) for c, n in zip(chroms, z["chrlen"])},
        genes={c: [str(g) for g in z[f"genes_{c}"]] for c in gchroms},
        gpos={c: [int(p) for p in z[f"gpos_{c}"]] for c in gchroms},
        tprob={c: z[f"tprob_{c}"] for c in gchroms},
        avecs={str(g): a for g, a in zip(z["avec_genes"], z["avecs"])},
        expr={str(g): v for g, v in zip(z["expr_genes"], z["expr"])}, truth={})


def main():
    for name, (gen_kw, rec_kw) in CASES.items():
        d = synth.generate_reconstruct(**gen_kw)
        out = pack_inputs(d)
        with tempfile.TemporaryDirectory() as tmp:
            p = synth.write_reconstruct_files(d, tmp)
            gu = load_reference_gbrs_utils(tmp)
            base = os.path.join(tmp, "ref")
            gu.reconstruct(expression_file=p["expr"], tprob_file=p["tprob"], avec_file=p["avecs"], gpos_file=p["gpos"],
                           outbase=base, **rec_kw)  # unmodified reference
            gp = np.load(base + ".genoprobs.npz")
            vs = np.load(base + ".genotypes.npz")
            for c in gp.files:
                out[f"gamma_{c}"] = gp[c]
            for c in vs.files:
                out[f"viterbi_{c}"] = vs[c]
            out["out_chroms"] = np.array(gp.files)
            out["genotypes_tsv"] = np.array(open(base + ".genotypes.tsv").read())
        out["expr_threshold"], out["sigma"] = rec_kw["expr_threshold"], rec_kw["sigma"]
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
        calls = out["genotypes_tsv"].item().count("\n") - 1
        agree = {c: float(np.mean(gp[c].argmax(axis=0) == d.truth[c])) for c in gp.files}
        print(name, "chromosomes", list(gp.files), "calls", calls, "posterior-argmax == simulated state:", agree)


if __name__ == "__main__":
    main()
