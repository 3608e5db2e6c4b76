"""Row-sharded path on ONE GPU: R shards packed separately, their local numerators summed by hand where the
cross-rank all-reduce would run, every shard then finishing the update from the same total -- exactly the sequence
`EMfactory.run` drives per rank (gbrs_em_launch_local -> exchange -> gbrs_em_launch_update).  The result must match the
unsharded run (identical iteration count, <= 1e-12 relative drift)."""
import ctypes as C
import os

import numpy as np
import pytest

from gbrs_b200 import _lib, synth
from gbrs_b200.emfactory import DevicePattern, EMfactory
from gbrs_b200.utils import gene_index
from tests import helpers as hp

pytestmark = pytest.mark.gpu


def run_sharded(d, model, R, tol=1e-4, max_iters=999):
    import torch

    apm = synth.to_apm(d)
    gene_of = gene_index(d.T, d.groups())
    eff = synth.effective_lengths(d)
    pats = [DevicePattern(apm, gene_of=gene_of, shard_rank=r, shard_count=R) for r in range(R)]
    assert sum(p.info["nnz"] for p in pats) == d.nnz
    lib = pats[0].lib
    for p in pats:
        p.set_lengths(eff)
        if model != 4:
            p.ensure_full()  # the arrays only models 1-3 read are uploaded on demand

    def exchange():
        total = torch.zeros_like(pats[0].acc)
        for p in pats:
            total += p.acc
        for p in pats:
            p.acc.copy_(total)

    for p in pats:
        _lib.check(lib.gbrs_em_prepare_local(C.byref(p.desc), p.stream()))
    exchange()
    for p in pats:
        _lib.check(lib.gbrs_em_prepare_finish(C.byref(p.desc), 0.0, p.stream()))
        _lib.check(lib.gbrs_em_run_begin(C.byref(p.desc), tol, max_iters, p.stream()))
    it = 0
    while True:
        for p in pats:
            _lib.check(lib.gbrs_em_launch_local(C.byref(p.desc), model, p.stream()))
        exchange()
        for p in pats:
            _lib.check(lib.gbrs_em_launch_update(C.byref(p.desc), p.stream()))
        it += 1
        ctrls = [p.read_ctrl()[0] for p in pats]
        assert len({int(c[_lib.CTRL_DONE]) for c in ctrls}) == 1  # every rank takes the same decision
        if ctrls[0][_lib.CTRL_DONE]:
            break
        assert it < max_iters + 2
    # two more (frozen) updates must not change anything: ranks over-run the stop by up to poll_every - 1 updates
    before = pats[0].current_theta_HT()
    for _ in range(2):
        for p in pats:
            _lib.check(lib.gbrs_em_launch_local(C.byref(p.desc), model, p.stream()))
        exchange()
        for p in pats:
            _lib.check(lib.gbrs_em_launch_update(C.byref(p.desc), p.stream()))
    thetas = [p.current_theta_HT() for p in pats]
    for th in thetas:
        assert np.array_equal(th, thetas[0])
    assert np.array_equal(before, thetas[0])
    return dict(theta=thetas[0], counts=pats[0].acc_HT(), iters=int(ctrls[0][_lib.CTRL_ITERS]))


@pytest.mark.parametrize("model", [4, 3, 2, 1])
@pytest.mark.parametrize("R", [2, 3])
def test_sharded_equals_unsharded(model, R, tmp_path):
    d = synth.generate(T=800, N=20000, H=8, sample_index=7)
    apm = synth.to_apm(d)
    em = EMfactory(apm)
    em.target_lengths = synth.effective_lengths(d)
    em.prepare()
    em.run(model=model, tol=1e-4, max_iters=999, verbose=False)
    s = run_sharded(d, model, R)
    assert s["iters"] == em.num_iters
    assert hp.relerr(s["theta"], em.allelic_expression) < 1e-12
    assert hp.relerr(s["counts"], em.expected_read_counts()) < 1e-12
    o = hp.oracle_run(d, model)
    assert s["iters"] == o["iters"] and hp.relerr(s["counts"], o["counts"]) < 1e-9
