"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/ by executing the unmodified reference
(/root/reference) in the build container.  Run:  python -m oracle.make_golden

Outputs (all small, committed):
  tests/golden/em_<name>.npz        inputs (pair form) + reference prepare()/run() outputs for
                                    models 1-4, multiway / diploid (-G) / pseudocount variants
  tests/golden/c1_model4.npz        config 1 (T=2000, N=200000, model 4): iteration count, exact
                                    err trajectory, theta, counts (inputs re-generated from the seed;
                                    an input checksum is stored)
  tests/golden/quantify_*/          the text tables written by the reference's own `quantify()`
                                    (gbrs/emase_utils.py:180-332) for a small multiway and diploid run

Provenance is recorded in tests/golden/PROVENANCE.json (numpy / scipy versions, shim use).
"""
from __future__ import annotations

import contextlib
import hashlib
import io
import json
import os
import shutil
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from gbrs_b200 import synth  # noqa: E402
from oracle import ref_harness as rh  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def input_checksum(d) -> str:
    m = hashlib.sha256()
    for a in (d.pair_class, d.pair_locus, d.pair_mask, d.count, d.gene_of, d.lengths):
        m.update(np.ascontiguousarray(a).tobytes())
    return m.hexdigest()


def traced_reference_run(d, model, lenfile, pseudocount, tol, max_iters, masked):
    """The reference's loop (EMfactory.py:264-279) driven from outside with the reference's own
    `update_allelic_expression`, only to capture err_sum at full precision (run() prints %9.1f)."""
    ref = rh.load_reference()
    rh.install_shim(model != 4)
    old = np.seterr(all="raise")
    np.seterr(under="ignore")
    try:
        apm = rh.build_reference_apm(d, masked=masked)
        em = ref.EMfactory(apm)
        em.prepare(pseudocount=pseudocount, lenfile=lenfile)
        errs = []
        err_sum, it = 1e6, 0
        while err_sum > 1e6 * tol and it < max_iters:
            prev = em.get_allelic_expression().sum(axis=0)
            prev *= 1e6 / prev.sum()
            em.update_allelic_expression(model=model)
            cur = em.get_allelic_expression().sum(axis=0)
            cur *= 1e6 / cur.sum()
            err_sum = float(np.abs(cur - prev).sum())
            errs.append(err_sum)
            it += 1
    finally:
        np.seterr(**old)
        rh.install_shim(False)
    return np.array(errs)


def em_case(name, T, N, H, model, masked, pseudocount, tol=1e-4, max_iters=999, store_inputs=True, n_genes=None):
    d = synth.generate(T=T, N=N, H=H, with_genotype=masked, n_genes=n_genes)
    with tempfile.TemporaryDirectory() as tmp:
        lenfile = os.path.join(tmp, "len.tsv")
        synth.write_length_file(d, lenfile)
        r = rh.run_reference_em(d, model, lenfile, pseudocount=pseudocount, tol=tol, max_iters=max_iters,
                                masked=masked)
        errs = traced_reference_run(d, model, lenfile, pseudocount, tol, max_iters, masked)
    assert len(errs) == r["iters"], (len(errs), r["iters"])
    out = dict(T=T, N=N, H=H, model=model, masked=masked, pseudocount=pseudocount, tol=tol, max_iters=max_iters,
               n_genes=-1 if n_genes is None else n_genes,
               iters=r["iters"], errs=errs, theta0=r["theta0"], theta=r["theta"], counts=r["counts"],
               checksum=input_checksum(d))
    if store_inputs:
        out.update(pair_class=d.pair_class.astype(np.int32), pair_locus=d.pair_locus.astype(np.int32),
                   pair_mask=d.pair_mask, count=d.count, gene_of=d.gene_of.astype(np.int32), lengths=d.lengths)
        if masked:
            out["gtmask"] = synth.genotype_mask(d)
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
    print(f"{name}: iters={r['iters']} last_err={errs[-1]:.6f} sum_counts={r['counts'].sum():.3f}")


def quantify_case(name, masked, model=4, pseudocount=0.0):
    """Unmodified reference `quantify()` end to end (SURVEY.md 8c recipe B)."""
    ref = rh.load_reference()
    rh.install_shim(model != 4)
    d = synth.generate(T=60, N=800, H=8, with_genotype=True)
    outdir = os.path.join(GOLD, name)
    shutil.rmtree(outdir, ignore_errors=True)
    os.makedirs(outdir)
    with tempfile.TemporaryDirectory() as tmp:
        apm = rh.build_reference_apm(d)
        h5 = os.path.join(tmp, "aln.h5")
        apm.save(h5file=h5)
        grp, ln, gt = (os.path.join(tmp, x) for x in ("grp.tsv", "len.tsv", "gt.tsv"))
        synth.write_group_file(d, grp)
        synth.write_length_file(d, ln)
        synth.write_genotype_file(d, gt)
        old = np.geterr()
        try:
            with contextlib.redirect_stdout(io.StringIO()) as so:
                ref.gutils.quantify(alignment_file=h5, group_file=grp, length_file=ln,
                                    genotype_file=gt if masked else None, outbase=os.path.join(tmp, "out"),
                                    multiread_model=model, pseudocount=pseudocount, max_iters=999, tolerance=1e-4,
                                    report_alignment_counts=True, report_posterior=False)
        finally:
            np.seterr(**old)
            rh.install_shim(False)
        with open(os.path.join(outdir, "stdout.txt"), "w") as fh:
            fh.write(so.getvalue())
        for fn in sorted(os.listdir(tmp)):
            if fn.startswith("out."):
                shutil.copy(os.path.join(tmp, fn), os.path.join(outdir, fn))
        for fn in ("grp.tsv", "len.tsv", "gt.tsv"):
            shutil.copy(os.path.join(tmp, fn), os.path.join(outdir, fn))
    np.savez_compressed(os.path.join(outdir, "input.npz"), T=d.T, H=d.H, N=d.N,
                        pair_class=d.pair_class.astype(np.int32), pair_locus=d.pair_locus.astype(np.int32),
                        pair_mask=d.pair_mask, count=d.count, model=model, pseudocount=pseudocount)
    print(name, sorted(os.listdir(outdir)))


def main():
    if not rh.reference_available():
        raise SystemExit("reference not mounted; golden vectors can only be generated in the build container")
    os.makedirs(GOLD, exist_ok=True)
    import scipy

    for model in (1, 2, 3, 4):
        em_case(f"em_small_m{model}", 200, 3000, 8, model, False, 0.0)
        em_case(f"em_small_m{model}_diploid", 200, 3000, 8, model, True, 0.0)
        em_case(f"em_small_m{model}_pc", 200, 3000, 8, model, False, 0.5)
    em_case("em_small_m4_h1", 150, 2000, 1, 4, False, 0.0)  # single haplotype: length file without _hap suffixes
    em_case("em_small_m3_h1", 150, 2000, 1, 3, False, 0.25)
    em_case("em_small_m4_h2", 150, 2000, 2, 4, False, 0.0)
    em_case("em_small_m2_h3", 150, 2000, 3, 2, False, 0.0)
    em_case("em_small_m4_maxit5", 200, 3000, 8, 4, False, 0.0, max_iters=5)
    em_case("em_small_m1_biggenes", 300, 4000, 8, 1, False, 0.0, n_genes=20)
    em_case("em_small_m3_biggenes", 300, 4000, 8, 3, False, 0.0, n_genes=20)
    em_case("c1_model4", 2000, 200000, 8, 4, False, 0.0, store_inputs=False)
    quantify_case("quantify_multiway", masked=False)
    quantify_case("quantify_diploid", masked=True)
    quantify_case("quantify_multiway_m2", masked=False, model=2, pseudocount=0.25)
    with open(os.path.join(GOLD, "PROVENANCE.json"), "w") as fh:
        json.dump({"reference": "churchill-lab/gbrs @ /root/reference (read-only mount)",
                   "numpy": np.__version__, "scipy": scipy.__version__, "python": sys.version.split()[0],
                   "generator": "python -m oracle.make_golden",
                   "note": "model 4 = unmodified reference; models 1-3 = reference control flow with the "
                           "sparse/sparse divide shim of oracle/ref_harness.py (reference crashes as-is "
                           "on this scipy, SURVEY.md fact 3); `tables` is a pickle-backed stand-in"}, fh, indent=1)


if __name__ == "__main__":
    main()
