"""Pin the oracle restatement (oracle/em_oracle.py) against golden vectors produced by the reference
itself (oracle/make_golden.py, executed in the build container)."""
import hashlib

import numpy as np
import pytest

from tests import helpers as hp


@pytest.mark.parametrize("name", hp.golden_em_cases())
def test_oracle_matches_reference_golden(name):
    g = hp.load_golden(name)
    d = hp.synth_from_golden(g)
    out = hp.oracle_run(d, g["model"], g["pseudocount"], g["tol"], g["max_iters"], g.get("gtmask"))
    assert out["iters"] == g["iters"]
    assert hp.relerr(out["theta0"], g["theta0"]) < 1e-12
    assert hp.relerr(out["theta"], g["theta"]) < 1e-12
    assert hp.relerr(out["counts"], g["counts"]) < 1e-12
    np.testing.assert_allclose(out["errs"], g["errs"], rtol=1e-9, atol=1e-9)
    # conservation: every class' posterior sums to one  =>  sum counts == sum class counts
    assert abs(out["counts"].sum() - g["counts"].sum()) < 1e-6


def test_oracle_config1_model4():
    g = hp.load_golden("c1_model4")
    d = hp.synth_from_golden(g)
    m = hashlib.sha256()
    for a in (d.pair_class, d.pair_locus, d.pair_mask, d.count, d.gene_of, d.lengths):
        m.update(np.ascontiguousarray(a).tobytes())
    if m.hexdigest() != str(g["checksum"]):
        pytest.skip("numpy RNG stream differs from the one the golden was generated with")
    out = hp.oracle_run(d, 4, 0.0, g["tol"], g["max_iters"])
    assert out["iters"] == g["iters"] == 17
    assert hp.relerr(out["theta"], g["theta"]) < 1e-12
    assert hp.relerr(out["counts"], g["counts"]) < 1e-12
    np.testing.assert_allclose(out["errs"], g["errs"], rtol=1e-9)
