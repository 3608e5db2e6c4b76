"""Checks shared by tests/test_reconstruct_simt.py (HMM kernels on the host SIMT shim) and
tests/test_zz_reconstruct_gpu.py (the same kernels on the GPU): a result dict (`gamma`, `states`, `eprob`, `alpha`,
`scaler`, `delta`, gene-major) for an `HmmPlan` against the oracle and against the reference's golden vectors."""
from __future__ import annotations

import os
from itertools import combinations_with_replacement

import numpy as np

from gbrs_b200 import reconstruct as rc
from oracle import reconstruct_oracle as ro
from oracle.make_golden_reconstruct import unpack_inputs

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["reconstruct_h8", "reconstruct_h2", "reconstruct_h4"]

# posterior / forward values: exp and log differ from numpy's by an ulp, sums run in another order -> 1e-10 relative
# (the bar for the EM is 1e-6); Viterbi scores, path and calls: bit-exact given the same emission values.
RTOL = 1e-10


def plan_of(d, tables=None):
    return rc.build_plan(d.chroms, d.genes, d.tprob, d.avecs, tables or [d.expr], d.H)


def genotype_names(d):
    return [a + b for a, b in combinations_with_replacement(d.hname, 2)]


def tsv_of(gtcall):
    return "#Gene_ID\tDiplotype\n" + "".join(f"{g}\t{gtcall[g]}\n" for g in sorted(gtcall))


def load_case(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return z, unpack_inputs(z), float(z["expr_threshold"]), float(z["sigma"])


def stress_case(H, sample_index=11):
    """Forbidden transitions (log 0 = -inf in 15 % of the off-diagonal entries) on top of the synthetic sample; run with
    a small sigma the emissions of most diplotypes underflow to the reference's log(nextafter(0, 1)) floor."""
    from gbrs_b200 import synth

    d = synth.generate_reconstruct(genes_per_chrom=(60, 9), H=H, sample_index=sample_index, extra_tprob_step=("X",))
    rng = np.random.default_rng(1)
    for c in d.tprob:
        t = d.tprob[c]
        S = t.shape[1]
        kill = rng.random(t.shape) < 0.15
        kill[:, np.arange(S), np.arange(S)] = False
        t[kill] = -np.inf
    return d


STRESS = [(8, 0.02), (2, 0.005), (8, 0.12)]  # (haplotypes, sigma)


def check_against_golden(plan, res, d, z):
    gamma, vit, gtcall = rc.collect_sample(plan, res, 0, genotype_names(d))
    assert sorted(gamma) == sorted(str(c) for c in z["out_chroms"])
    for c in gamma:
        assert gamma[c].shape == z[f"gamma_{c}"].shape
        np.testing.assert_allclose(gamma[c], z[f"gamma_{c}"], rtol=RTOL, atol=1e-300)
        assert vit[c] == list(z[f"viterbi_{c}"])
    assert tsv_of(gtcall) == z["genotypes_tsv"].item()


def check_against_oracle(plan, res, d, tables, thr, sigma):
    genotypes = genotype_names(d)
    for s, table in enumerate(tables):
        want = ro.reconstruct_tables(d.chroms, d.genes, d.tprob, d.avecs, table, d.hname, thr, sigma)
        gamma, vit, gtcall = rc.collect_sample(plan, res, s, genotypes)
        assert sorted(gamma) == sorted(want["gamma"])
        own_eprob = {}
        for ci, c in enumerate(plan.chroms):
            ch = plan.chain_of(s, ci)
            g0, n = int(ch["gene0"]), int(ch["n_genes"])
            e_want = np.array([want["eprob"][g] for g in d.genes[c]])
            # emissions: exp() results below ~1e-304 are denormals whose last bits depend on the exp implementation;
            # compared above that, and both sides must agree on "numerically impossible" below
            e_got = res["eprob"][g0:g0 + n]
            normal = e_want > -700.0
            np.testing.assert_allclose(e_got[normal], e_want[normal], rtol=RTOL, atol=1e-12)
            assert (e_got[~normal] < -690.0).all() and np.isfinite(e_got).all()
            det = want["detail"][c]
            # forward log-probabilities: compared where the state has a representable probability.  Below ~1e-280 the
            # terms are denormals (a few bits) or hit the reference's `+ nextafter(0, 1)` floor; the kernel forms them
            # as products exp(alpha) * exp(tprob) instead of exp(alpha + tprob) and rounds differently there -- both
            # sides must agree that the state is (numerically) impossible, and the posterior check below still applies
            got_alpha = res["alpha"][g0:g0 + n].T
            live = det["alpha"] > -640.0
            np.testing.assert_allclose(got_alpha[live], det["alpha"][live], rtol=RTOL, atol=1e-10)
            assert (got_alpha[~live] < -600.0).all() and np.isfinite(got_alpha).all()
            np.testing.assert_allclose(res["scaler"][g0:g0 + n], det["scaler"], rtol=RTOL, atol=1e-10)
            np.testing.assert_allclose(gamma[c], det["gamma"], rtol=RTOL, atol=1e-300)
            np.testing.assert_allclose(gamma[c].sum(axis=0), 1.0, rtol=1e-12)
            for i, g in enumerate(d.genes[c]):
                own_eprob[g] = res["eprob"][g0 + i]
        for g in table:
            own_eprob.setdefault(g, want["eprob"][g])
        # the chain arithmetic of the Viterbi part is additions and comparisons only: on the kernel's own emission
        # values the oracle must reproduce scores, path and calls exactly
        exact = ro.reconstruct_tables(d.chroms, d.genes, d.tprob, d.avecs, table, d.hname, thr, sigma, eprob=own_eprob)
        for ci, c in enumerate(plan.chroms):
            ch = plan.chain_of(s, ci)
            g0, n = int(ch["gene0"]), int(ch["n_genes"])
            assert np.array_equal(res["delta"][g0:g0 + n].T, exact["detail"][c]["delta"])
            assert vit[c] == exact["viterbi"][c]
        assert gtcall == exact["gtcall"]
