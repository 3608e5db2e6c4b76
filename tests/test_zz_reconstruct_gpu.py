"""`reconstruct` on the GPU through the C ABI (gbrs_hmm_emission / gbrs_hmm_run) and through the workflow function,
against the oracle and the golden vectors written by the unmodified reference.  The same checks run against the same
kernel source on the CPU in tests/test_reconstruct_simt.py.

The file name sorts last on purpose: these kernels were written after the round's GPU budget was spent, so the first run
on a real device is the driver's; the parity tests of the EM path come first."""
import os

import numpy as np
import pytest

from gbrs_b200 import reconstruct as rc
from gbrs_b200 import synth
from tests import reconstruct_checks as chk

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", chk.CASES)
def test_gpu_kernels_match_reference_golden(name):
    z, d, thr, sigma = chk.load_case(name)
    plan = chk.plan_of(d)
    res = rc.run_plan_on_device(plan, thr, sigma, keep_work=True)
    chk.check_against_golden(plan, res, d, z)
    chk.check_against_oracle(plan, res, d, [d.expr], thr, sigma)


def test_gpu_cohort_launch_and_long_chain():
    kw = dict(genes_per_chrom=(700, 7, 1, 2), H=8, extra_tprob_step=("2",))
    base = synth.generate_reconstruct(sample_index=0, **kw)
    tables = [base.expr] + [synth.generate_reconstruct(sample_index=s, **kw).expr for s in (1, 2, 3)]
    plan = chk.plan_of(base, tables=tables)
    res = rc.run_plan_on_device(plan, 1.5, 0.12, keep_work=True)
    chk.check_against_oracle(plan, res, base, tables, 1.5, 0.12)
    again = rc.run_plan_on_device(plan, 1.5, 0.12, keep_work=True)  # no atomics, fixed summation order
    for k in ("gamma", "states", "eprob", "alpha", "delta"):
        assert np.array_equal(res[k], again[k])


@pytest.mark.parametrize("H", [1, 3, 5, 6, 7])
def test_gpu_other_haplotype_counts(H):
    d = synth.generate_reconstruct(genes_per_chrom=(40, 9), H=H, sample_index=H)
    plan = chk.plan_of(d)
    res = rc.run_plan_on_device(plan, 1.0, 0.15, keep_work=True)
    chk.check_against_oracle(plan, res, d, [d.expr], 1.0, 0.15)


@pytest.mark.parametrize("H,sigma", chk.STRESS)
def test_gpu_forbidden_transitions_and_underflowing_emissions(H, sigma):
    d = chk.stress_case(H)
    plan = chk.plan_of(d)
    res = rc.run_plan_on_device(plan, 1.5, sigma, keep_work=True)
    assert np.isfinite(res["gamma"]).all()
    with np.errstate(all="ignore"):
        chk.check_against_oracle(plan, res, d, [d.expr], 1.5, sigma)


def test_gpu_reconstruct_files_match_reference(tmp_path, monkeypatch):
    """File to file through `reconstruct()`: the three output files against the reference's."""
    z, d, thr, sigma = chk.load_case("reconstruct_h8")
    p = synth.write_reconstruct_files(d, str(tmp_path))
    monkeypatch.setenv("GBRS_DATA", p["data_dir"])
    base = str(tmp_path / "out")
    rc.reconstruct(expression_file=p["expr"], tprob_file=p["tprob"], avec_file=p["avecs"], gpos_file=p["gpos"],
                   expr_threshold=thr, sigma=sigma, outbase=base)
    gp = np.load(base + ".genoprobs.npz")
    vs = np.load(base + ".genotypes.npz")
    assert sorted(gp.files) == sorted(str(c) for c in z["out_chroms"])
    for c in gp.files:
        np.testing.assert_allclose(gp[c], z[f"gamma_{c}"], rtol=chk.RTOL, atol=1e-300)
        assert list(vs[c]) == list(z[f"viterbi_{c}"])
    assert open(base + ".genotypes.tsv").read() == z["genotypes_tsv"].item()
    # default file locations under $GBRS_DATA (gbrs_utils.py:408-412) and default output names (:399-401)
    os.replace(p["avecs"], os.path.join(p["data_dir"], "avecs.npz"))
    monkeypatch.chdir(tmp_path)
    rc.reconstruct(expression_file=p["expr"], tprob_file=p["tprob"], expr_threshold=thr, sigma=sigma)
    assert open("gbrs.reconstructed.genotypes.tsv").read() == z["genotypes_tsv"].item()
    assert os.path.exists("gbrs.reconstructed.genoprobs.npz") and os.path.exists("gbrs.reconstructed.genotypes.npz")
