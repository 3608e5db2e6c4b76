"""The HMM kernels of `reconstruct` (gbrs_b200/csrc/hmm_kernels.cu) executed on the CPU through the host SIMT shim
(tests/simt/: the same kernel source, one OS thread per CUDA thread, pthread barriers, deferred cp.async), checked
against the oracle and against the golden vectors of the unmodified reference.  This is a check of the kernel CODE on
machines without a GPU -- not a product path: the package cannot reach the emulation library.  The same comparisons run
against the real kernels in tests/test_zz_reconstruct_gpu.py."""
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest

from gbrs_b200 import reconstruct as rc
from gbrs_b200 import synth
from tests import reconstruct_checks as chk
from tests import simt_emul

pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="g++ is needed to build the SIMT emulation")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("name", chk.CASES)
def test_emulated_kernels_match_reference_golden(name):
    z, d, thr, sigma = chk.load_case(name)
    plan = chk.plan_of(d)
    res = simt_emul.run_plan_emulated(plan, thr, sigma)
    chk.check_against_golden(plan, res, d, z)
    chk.check_against_oracle(plan, res, d, [d.expr], thr, sigma)


def test_emulated_cohort_and_grid_stride():
    """Three samples in one launch (chains share transition matrices), fewer blocks than chains (grid-stride loop,
    shared memory reused by the next chain), a chromosome longer than one back-trace round (256 genes)."""
    kw = dict(genes_per_chrom=(300, 7, 1, 2), H=3, extra_tprob_step=("2",))
    base = synth.generate_reconstruct(sample_index=0, **kw)
    tables = [base.expr] + [synth.generate_reconstruct(sample_index=s, **kw).expr for s in (1, 2)]
    plan = chk.plan_of(base, tables=tables)
    assert len(plan.chains) == 12 and plan.n_samples == 3
    res = simt_emul.run_plan_emulated(plan, 1.5, 0.12, grid=5)
    chk.check_against_oracle(plan, res, base, tables, 1.5, 0.12)


@pytest.mark.parametrize("H", [1, 5, 6, 7])
def test_emulated_other_haplotype_counts(H):
    d = synth.generate_reconstruct(genes_per_chrom=(12, 9), H=H, sample_index=H)
    plan = chk.plan_of(d)
    res = simt_emul.run_plan_emulated(plan, 1.0, 0.15)
    chk.check_against_oracle(plan, res, d, [d.expr], 1.0, 0.15)


@pytest.mark.parametrize("H,sigma", chk.STRESS)
def test_emulated_forbidden_transitions_and_underflowing_emissions(H, sigma):
    d = chk.stress_case(H)
    plan = chk.plan_of(d)
    with np.errstate(all="ignore"):
        res = simt_emul.run_plan_emulated(plan, 1.5, sigma)
        assert res["eprob"].min() < -700.0 or sigma > 0.1  # the floor is really reached
        assert np.isfinite(res["gamma"]).all()
        chk.check_against_oracle(plan, res, d, [d.expr], 1.5, sigma)


def test_kernels_are_race_free_under_thread_sanitizer():
    """The emulation built with -fsanitize=thread: every pair of shared / global memory accesses of the kernels that
    is not ordered by a barrier is reported.  Runs in a subprocess (the sanitizer runtime has to be loaded first)."""
    try:
        lib = simt_emul.build(tsan=True)
    except RuntimeError as e:
        pytest.skip(f"ThreadSanitizer build unavailable: {e}")
    rt = subprocess.run(["g++", "-print-file-name=libtsan.so"], capture_output=True, text=True).stdout.strip()
    if not os.path.isabs(rt) or not os.path.exists(rt):
        pytest.skip("libtsan.so not found")
    code = f"""
import sys
sys.path.insert(0, {ROOT!r})
import numpy as np
from gbrs_b200 import reconstruct as rc, synth
from tests import simt_emul
lib = simt_emul.bind({lib!r})
d = synth.generate_reconstruct(genes_per_chrom=(270, 6, 1), H=4, sample_index=1, extra_tprob_step=("2",))
plan = rc.build_plan(d.chroms, d.genes, d.tprob, d.avecs, [d.expr, d.expr], d.H)
e = simt_emul.run_emission(plan, 1.5, 0.12, lib=lib)
r = simt_emul.run_chains(plan, e, grid=2, lib=lib)
assert np.isfinite(r["gamma"]).all()
print("TSAN-RUN-OK")
"""
    env = dict(os.environ, LD_PRELOAD=rt, TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0 exitcode=0")
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=900)
    if "TSAN-RUN-OK" not in res.stdout:
        pytest.skip("the sanitizer run did not complete here: " + res.stderr[-400:])
    assert "ThreadSanitizer: data race" not in res.stderr, res.stderr[-3000:]


def test_workflow_files_with_emulated_kernels(tmp_path, monkeypatch):
    """`reconstruct()` file to file -- parsing, chain layout, output files -- with the device step replaced by the
    emulated kernels (a test-only substitution): the three output files against the unmodified reference's."""
    z, d, thr, sigma = chk.load_case("reconstruct_h8")
    p = synth.write_reconstruct_files(d, str(tmp_path))
    monkeypatch.setenv("GBRS_DATA", p["data_dir"])
    monkeypatch.setattr(rc, "run_plan_on_device",
                        lambda plan, t, s, device=None, keep_work=False: simt_emul.run_plan_emulated(plan, t, s))
    base = str(tmp_path / "out")
    rc.reconstruct(expression_file=p["expr"], tprob_file=p["tprob"], avec_file=p["avecs"], gpos_file=p["gpos"],
                   expr_threshold=thr, sigma=sigma, outbase=base)
    gp, vs = np.load(base + ".genoprobs.npz"), np.load(base + ".genotypes.npz")
    assert sorted(gp.files) == sorted(str(c) for c in z["out_chroms"])
    for c in gp.files:
        np.testing.assert_allclose(gp[c], z[f"gamma_{c}"], rtol=chk.RTOL, atol=1e-300)
        assert list(vs[c]) == list(z[f"viterbi_{c}"])
    assert open(base + ".genotypes.tsv").read() == z["genotypes_tsv"].item()
    # cohort form: two samples, one launch, one set of files per sample
    other = synth.generate_reconstruct(genes_per_chrom=(22, 13, 1, 16), H=8, sample_index=9, extra_tprob_step=("2",))
    other.genes, other.tprob, other.avecs = d.genes, d.tprob, d.avecs
    p2 = synth.write_reconstruct_files(other, str(tmp_path), prefix="s2.")
    rc.reconstruct_cohort([p["expr"], p2["expr"]], p["tprob"], avec_file=p["avecs"], gpos_file=p["gpos"],
                          expr_threshold=thr, sigma=sigma, outbases=[str(tmp_path / "a"), str(tmp_path / "b")])
    assert open(str(tmp_path / "a") + ".genotypes.tsv").read() == z["genotypes_tsv"].item()
    from oracle import reconstruct_oracle as ro
    want = ro.reconstruct_tables(d.chroms, d.genes, d.tprob, d.avecs, rc.read_expression(p2["expr"])[1], d.hname, thr,
                                 sigma)
    assert open(str(tmp_path / "b") + ".genotypes.tsv").read() == chk.tsv_of(want["gtcall"])
