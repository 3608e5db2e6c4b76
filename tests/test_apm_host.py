"""Host-container methods of `AlignmentPropertyMatrix` that callers of the reference use around the EM: construction
by shape, `copy`, `finalize`, `reset`, the `-G` form of `multiply`, `_bundle_inline` (AlignmentPropertyMatrix.py:25-111,
:132-188; Sparse3DMatrix.py:220-234, :354-362) -- against direct expectations and, where the reference sources are
present, against the reference's own container."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

from gbrs_b200 import synth
from gbrs_b200.apm import AlignmentPropertyMatrix as APM


@pytest.fixture()
def data():
    return synth.generate(T=40, N=300, H=4, with_genotype=True)


def dense(m):
    return np.asarray(sp.csc_matrix(m).todense())


def test_construct_by_shape_fill_finalize(data):
    d = data
    apm = APM(shape=(d.T, d.H, d.N), haplotype_names=list(d.hname), locus_names=d.lname)
    assert (apm.num_loci, apm.num_haplotypes, apm.num_reads) == (d.T, d.H, d.N) and not apm.finalized
    assert apm.lid[d.lname[3]] == 3 and apm.count is None
    mats = synth.to_csc_list(d)
    for h in range(d.H):
        apm.data[h] = mats[h].tocoo()
    with pytest.raises(RuntimeError, match="finalized"):
        apm.multiply(np.ones((d.H, d.T)), axis=2)
    apm.finalize()
    assert apm.finalized and all(m.format == "csc" for m in apm.data) and apm.nnz == d.nnz
    for bad in (dict(shape=(d.T, d.H)), dict(shape=(d.T, 0, d.N))):
        with pytest.raises(RuntimeError):
            APM(**bad)
    with pytest.raises(RuntimeError, match="number of names"):
        APM(shape=(d.T, d.H, d.N), haplotype_names=["A"])
    with pytest.raises(RuntimeError, match="number of names"):
        APM(shape=(d.T, d.H, d.N), locus_names=d.lname[:-1])


def test_copy_is_deep_and_shallow_drops_names(data):
    apm = synth.to_apm(data)
    c = apm.copy()
    assert c.shape == apm.shape and c.finalized and list(c.lname) == list(apm.lname) and c.hname == apm.hname
    assert c.num_groups == apm.num_groups and c.groups == apm.groups and c.groups is not apm.groups
    c.data[0].data[:] = 7.0
    c.count[0] += 1
    assert apm.data[0].data.max() == 1.0 and apm.count[0] == data.count[0]
    s = apm.copy(shallow=True)
    assert s.lname is None and s.hname is None and s.groups is None and np.array_equal(s.count, apm.count)
    with pytest.raises(RuntimeError, match="finalized"):
        APM(other=APM(shape=(3, 2, 4)))


def test_multiply_by_genotype_mask_and_reset(data):
    d = data
    apm = synth.to_apm(d)
    gm = synth.genotype_mask(d)
    before = [dense(m) for m in apm.data]
    apm.multiply(gm, axis=2)
    for h in range(d.H):
        assert np.array_equal(dense(apm.data[h]), before[h] * gm[h][None, :])
    assert not apm.is_pure_incidence()  # explicit zeros until eliminate_zeros()
    apm.eliminate_zeros()
    assert apm.is_pure_incidence() and apm.nnz == int(sum((before[h] * gm[h][None, :]).sum() for h in range(d.H)))
    apm.data[1].data[:] = 0.25
    apm.reset()
    assert apm.is_pure_incidence()
    with pytest.raises(NotImplementedError):
        apm.multiply(np.ones(d.T), axis=1)


def test_bundle_inline_sums_loci_into_genes(data):
    d = data
    apm = synth.to_apm(d)
    before = [dense(m) for m in apm.data]
    groups, gname = apm.groups, list(apm.gname)
    apm._bundle_inline(reset=False)
    G = len(gname)
    assert apm.shape == (G, d.H, d.N) and apm.num_loci == G and list(apm.lname) == gname and apm.lid[gname[2]] == 2
    assert apm.num_groups == 0 and apm.groups is None and apm.gname is None
    for h in range(d.H):
        want = np.stack([before[h][:, groups[g]].sum(axis=1) for g in range(G)], axis=1)
        assert np.array_equal(dense(apm.data[h]), want)
    assert max(m.data.max() for m in apm.data) > 1.0  # a class hitting two isoforms of a gene counts twice ...
    again = synth.to_apm(d)
    again._bundle_inline(reset=True)
    assert again.is_pure_incidence()  # ... unless reset: gene-level incidence
    with pytest.raises(RuntimeError, match="No group information"):
        apm._bundle_inline()


@pytest.mark.skipif(not os.path.isdir("/root/reference/src/gbrs/emase"), reason="needs the reference sources")
def test_host_container_equals_the_reference_container(data, monkeypatch):
    import sys

    from oracle import ref_harness as rh

    monkeypatch.setitem(sys.modules, "tables", rh._fake_tables_module())  # do not leave the stand-in behind
    rh.load_reference()
    d = data
    ours, theirs = synth.to_apm(d), rh.build_reference_apm(d)
    gm = synth.genotype_mask(d)
    for a in (ours, theirs):
        a.multiply(gm, axis=2)
    for h in range(d.H):
        theirs.data[h].eliminate_zeros()
    ours.eliminate_zeros()
    for h in range(d.H):
        assert (sp.csc_matrix(ours.data[h]) != sp.csc_matrix(theirs.data[h])).nnz == 0
    co, ct = ours.copy(), theirs.copy()
    for a in (co, ct):
        a._bundle_inline(reset=True)
    assert tuple(co.shape) == tuple(ct.shape) and list(co.lname) == list(ct.lname)
    for h in range(d.H):
        assert (sp.csc_matrix(co.data[h]) != sp.csc_matrix(ct.data[h])).nnz == 0


def test_group_tables_follow_the_reference_loop():
    """`grp_conv_mat[groups[i], i] = 1.0` (EMfactory.py:42-47) and the gene id per locus the kernels use instead of
    `t2t_mat`; loci outside every group get ids of their own past the real genes."""
    from gbrs_b200 import utils

    groups = [[0, 3], [5], [], [1, 2, 2]]  # an empty group and a repeated locus
    T = 8
    want = np.zeros((T, len(groups)))
    for i, g in enumerate(groups):
        want[g, i] = 1.0
    m = utils.group_conversion_matrix(T, groups)
    assert m.format == "csc" and np.array_equal(m.toarray(), want)
    g = utils.gene_index(T, groups)
    assert g.dtype == np.int32 and list(g[[0, 3, 5, 1, 2]]) == [0, 0, 1, 3, 3]
    assert sorted(g[[4, 6, 7]]) == [4, 5, 6] and len(set(g[[4, 6, 7]])) == 3
    with pytest.raises(NotImplementedError, match="more than one gene"):
        utils.gene_index(T, [[0, 1], [1, 2]])
    log = utils.configure_logging("gbrs-test", 2)
    assert log.level == 10 and utils.configure_logging("gbrs-test", 1).level == 20 and len(log.handlers) == 1
    assert utils.configure_logging("gbrs-test", 0).level == 30 and not log.propagate
