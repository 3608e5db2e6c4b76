"""The CPU restatement of `reconstruct` (oracle/reconstruct_oracle.py) against golden vectors written by the unmodified
reference (oracle/make_golden_reconstruct.py): posterior per chromosome, ordered Viterbi states, genotype table."""
import os

import numpy as np
import pytest

from oracle import reconstruct_oracle as ro
from oracle.make_golden_reconstruct import unpack_inputs

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
CASES = ["reconstruct_h8", "reconstruct_h2", "reconstruct_h4"]


def run_oracle(z):
    d = unpack_inputs(z)
    return d, ro.reconstruct_tables(d.chroms, d.genes, d.tprob, d.avecs, d.expr, d.hname,
                                    expr_threshold=float(z["expr_threshold"]), sigma=float(z["sigma"]))


def tsv_of(gtcall):
    return "#Gene_ID\tDiplotype\n" + "".join(f"{g}\t{gtcall[g]}\n" for g in sorted(gtcall))


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_reconstruct(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    d, got = run_oracle(z)
    assert sorted(got["gamma"]) == sorted(str(c) for c in z["out_chroms"])
    for c in got["gamma"]:
        want = z[f"gamma_{c}"]
        assert got["gamma"][c].shape == want.shape == (d.S, len(d.genes[c]))
        np.testing.assert_allclose(got["gamma"][c], want, rtol=1e-12, atol=1e-300)
        assert list(z[f"viterbi_{c}"]) == got["viterbi"][c]
    assert tsv_of(got["gtcall"]) == z["genotypes_tsv"].item()


def test_golden_set_covers_the_special_cases():
    z = np.load(os.path.join(GOLDEN, "reconstruct_h8.npz"))
    d, got = run_oracle(z)
    n = {c: len(d.genes[c]) for c in d.genes}
    steps = {c: len(d.tprob[c]) for c in d.genes}
    assert any(n[c] == 1 for c in n)  # one-gene chromosome: a Viterbi state, no genotype call
    assert any(steps[c] == n[c] and n[c] > 1 for c in n)  # legacy transition file with one matrix per gene
    assert any(steps[c] == n[c] - 1 and n[c] > 1 for c in n)
    assert any(c not in d.tprob for c in d.chroms)  # chromosome of the fai file without data: skipped
    init = ro.initial_logprob(d.H)
    kinds = {"null": 0, "naive": 0, "avec": 0}
    for g, v in d.expr.items():
        if sum(v) < float(z["expr_threshold"]):
            kinds["null"] += 1
            assert np.array_equal(got["eprob"][g], init)
        elif g not in d.avecs:
            kinds["naive"] += 1
        else:
            kinds["avec"] += 1
    assert all(k > 0 for k in kinds.values()), kinds
    # the last gene of a chromosome with genes - 1 matrices gets a Viterbi state but no call (gbrs_utils.py:585-594)
    for c in got["viterbi"]:
        called = got["detail"][c]["called"]
        assert len(got["viterbi"][c]) == called + 1
        if steps[c] == n[c] - 1:
            assert d.genes[c][-1] not in got["gtcall"]


def test_posterior_columns_sum_to_one_and_forward_is_normalised():
    z = np.load(os.path.join(GOLDEN, "reconstruct_h4.npz"))
    d, got = run_oracle(z)
    for c, r in got["detail"].items():
        np.testing.assert_allclose(r["gamma"].sum(axis=0), 1.0, rtol=1e-13)
        np.testing.assert_allclose(np.exp(r["alpha"]).sum(axis=0), 1.0, rtol=1e-12)
