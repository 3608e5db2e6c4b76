"""The TILE layout of the fused model-4 update (include/gbrs_em.h; gbrs_b200/csrc/tile_pack.cpp builds it, k_tile_em in
gbrs_b200/csrc/em_kernels.cu reads it).  CPU side: (1) a numpy walk over the blobs (tests/tile_emulation.py) must
reproduce the oracle's E-step + count-weighted column sums -- this pins every array of the format; (2) the kernel source
itself, executed through the host SIMT shim, must reproduce the reference's golden vectors.  The GPU tests
(tests/test_em_gpu.py and friends) run the same path on the device: a pattern uses the tile layout by default."""
import os
import shutil

import numpy as np
import pytest

from gbrs_b200 import synth
from gbrs_b200.emfactory import PackedPattern, TiledPattern
from gbrs_b200.quantify import hapmask_bytes
from oracle import em_oracle as eo
from tests import helpers as hp
from tests import tile_emulation as te

SMALL_CAPS = dict(max_classes=256, max_loci=32, max_pairs=1024, max_entries=1536, max_items=512)
TINY_CAPS = dict(max_classes=40, max_loci=24, max_pairs=96, max_entries=160, max_items=64, item_len=4)


def theta_T8(theta_HT, T):
    out = np.zeros((T, 8))
    out[:, : theta_HT.shape[0]] = theta_HT.T
    return out


@pytest.mark.parametrize("shape", [
    dict(T=400, N=8000, H=8), dict(T=120, N=1500, H=8, wide_frac=0.08, caps=SMALL_CAPS),
    dict(T=300, N=4000, H=3), dict(T=50, N=300, H=1), dict(T=90, N=1200, H=5, caps=TINY_CAPS),
    dict(T=200, N=3000, H=8, diploid=True, caps=TINY_CAPS)])
def test_tile_layout_reproduces_the_oracle_update(shape):
    shape = dict(shape)
    caps, diploid = shape.pop("caps", SMALL_CAPS), shape.pop("diploid", False)
    d = synth.generate(sample_index=2, with_genotype=diploid, **shape)
    T, H = d.T, d.H
    oapm = eo.apm_from_pairs(d.T, d.H, d.N, d.pair_class, d.pair_locus, d.pair_mask, d.count)
    hm = None
    if diploid:
        gm = synth.genotype_mask(d)
        oapm = eo.apply_genotype_mask(oapm, gm)
        hm = hapmask_bytes(gm)
    packed = PackedPattern(synth.to_apm(d), hapmask=hm)
    tiled = TiledPattern(packed, **caps)
    i = tiled.info
    assert i["n_classes"] == packed.info["n_classes"] and i["n_pairs"] == packed.info["n_pairs"]
    assert i["max_classes"] <= caps["max_classes"] and i["max_loci"] <= caps["max_loci"]
    assert i["max_items"] <= caps["max_items"] and i["n_tiles"] >= 1
    eff = eo.effective_length_table(d.lengths)
    theta = eo.prepare(oapm, eff, 0.0)
    # prepare(): theta0 * efflen = sum_n count / nnz
    W0 = te.numerator_W(tiled, None, T, unit=True)
    assert hp.relerr((W0[:, :H].T) / eff, theta) < 1e-13
    # one model-4 update: numerator = theta * W
    th8 = theta_T8(theta, T)
    acc = (th8 * te.numerator_W(tiled, th8, T))[:, :H].T
    want = eo.sum_read(oapm, eo.e_step(oapm, theta, 4, None))
    assert hp.relerr(acc, want) < 1e-13
    assert abs(acc.sum() - want.sum()) < 1e-10 * want.sum()


def test_tile_layout_is_independent_of_the_thread_count(monkeypatch):
    d = synth.generate(T=150, N=2500, H=8, sample_index=4)
    apm = synth.to_apm(d)
    blobs = []
    for nt in ("1", "3", "7"):
        monkeypatch.setenv("GBRS_PACK_THREADS", nt)
        tiled = TiledPattern(PackedPattern(apm), **TINY_CAPS)
        blobs.append({k: v.copy() for k, v in tiled.arrays.items()})
    for b in blobs[1:]:
        assert all(np.array_equal(b[k], blobs[0][k]) for k in b)


def test_class_wider_than_a_tile_is_refused():
    """A class touching more loci than a tile may hold: GBRS_E_LIMIT -> NotImplementedError; DevicePattern then keeps
    model 4 on the two-pass kernels."""
    d = synth.generate(T=200, N=500, H=8, sample_index=1, wide_frac=0.2, wide_max=40)
    assert np.bincount(d.pair_class).max() > 24
    packed = PackedPattern(synth.to_apm(d))
    with pytest.raises(NotImplementedError):
        TiledPattern(packed, **TINY_CAPS)
    with pytest.raises(Exception):
        TiledPattern(packed, max_loci=500)  # caps out of range


def test_empty_shard_has_no_tiles():
    from scipy.sparse import csc_matrix

    from gbrs_b200.apm import AlignmentPropertyMatrix

    apm = AlignmentPropertyMatrix.from_csc([csc_matrix((5, 7)) for _ in range(2)], ["A", "B"], [f"t{i}" for i in range(7)])
    tiled = TiledPattern(PackedPattern(apm))
    assert tiled.info["n_tiles"] == 0 and tiled.info["n_slots"] == 0
    ld = tiled.arrays["locus_desc"].reshape(-1, 4)
    assert ld.shape[0] == 7 and (ld[:, 1] == ld[:, 2]).all()


# ---- the kernel source on the CPU (host SIMT shim) -------------------------------------------------------------------
needs_gxx = pytest.mark.skipif(shutil.which("g++") is None, reason="g++ is needed to build the SIMT emulation")
GOLDEN_M4 = ["em_small_m4", "em_small_m4_diploid", "em_small_m4_h1"]
if os.environ.get("GBRS_SIMT_ALL"):
    GOLDEN_M4 = [n for n in hp.golden_em_cases() if n.startswith("em_small_m4")]


@needs_gxx
@pytest.mark.parametrize("name", GOLDEN_M4)
def test_emulated_tile_kernel_matches_reference_golden(name):
    from tests import simt_em

    g = hp.load_golden(name)
    d = hp.synth_from_golden(g)
    hm = hapmask_bytes(g["gtmask"]) if g["masked"] else None
    pat = simt_em.HostPattern(synth.to_apm(d), gene_of=eo.gene_index(d.T, d.groups()), hapmask=hm, tiles=TINY_CAPS)
    assert pat.tiled.info["n_tiles"] > 3
    theta0 = pat.prepare(eo.effective_length_table(d.lengths), g["pseudocount"])
    assert hp.relerr(theta0, g["theta0"]) < 1e-12
    out = pat.run(g["model"], g["tol"], g["max_iters"])
    assert out["iters"] == g["iters"]
    np.testing.assert_allclose(out["errs"], g["errs"], rtol=1e-7, atol=1e-7)
    assert hp.relerr(out["theta"], g["theta"]) < 1e-12 and hp.relerr(out["counts"], g["counts"]) < 1e-12


@needs_gxx
def test_emulated_tile_kernel_wide_classes_then_models_1_to_3():
    """Wide classes (many pair planes), several blocks pulling tiles from the work counter, and the two-pass kernels of
    models 1-3 taking over on the same pattern (the subset tables the tile kernel reads are shared with them)."""
    from tests import simt_em

    d = synth.generate(T=120, N=1500, H=8, sample_index=9, wide_frac=0.08)
    gene_of = eo.gene_index(d.T, d.groups())
    pat = simt_em.HostPattern(synth.to_apm(d), gene_of=gene_of, tiles=SMALL_CAPS)
    assert pat.tiled.info["max_planes"] > 8
    eff = eo.effective_length_table(d.lengths)
    theta = pat.prepare(eff)
    oapm = eo.apm_from_pairs(d.T, d.H, d.N, d.pair_class, d.pair_locus, d.pair_mask, d.count)
    assert hp.relerr(theta, eo.prepare(oapm, eff, 0.0)) < 1e-12
    for model in (4, 3, 4, 2, 1, 4):
        want = eo.sum_read(oapm, eo.e_step(oapm, theta, model, gene_of))
        out = pat.run(model, tol=0.0, max_iters=1)
        assert out["iters"] == 1 and hp.relerr(out["counts"], want) < 1e-12
        theta = out["theta"]
        assert hp.relerr(theta, want / eff) < 1e-12
