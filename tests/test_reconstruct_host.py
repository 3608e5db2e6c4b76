"""Host side of `reconstruct` (file parsing, chain layout, error behaviour) -- no GPU needed."""
import numpy as np
import pytest

from gbrs_b200 import _lib
from gbrs_b200 import reconstruct as rc
from gbrs_b200 import synth
from oracle import reconstruct_oracle as ro


@pytest.fixture()
def files(tmp_path):
    d = synth.generate_reconstruct(genes_per_chrom=(9, 5, 1), H=8, sample_index=2, extra_tprob_step=("2",))
    return d, synth.write_reconstruct_files(d, str(tmp_path))


def test_file_readers_follow_the_reference_formats(files):
    d, p = files
    chrlens = rc.get_chromosome_info(p["data_dir"])
    assert list(chrlens) == d.chroms and chrlens["MT"] == 16299
    haps, expr = rc.read_expression(p["expr"])
    assert haps == list(d.hname)
    assert list(expr) == list(d.expr)
    for g in d.expr:
        assert np.array_equal(expr[g], d.expr[g])  # repr() round trip
    order = rc.read_gene_order(p["gpos"])
    assert {c: list(v) for c, v in order.items()} == {c: d.genes[c] for c in d.genes}
    with pytest.raises(ValueError, match="ref.fa.fai"):
        rc.get_chromosome_info(p["data_dir"] + "/nowhere")


def test_plan_layout(files):
    d, p = files
    other = synth.generate_reconstruct(genes_per_chrom=(9, 5, 1), H=8, sample_index=3, extra_tprob_step=("2",))
    plan = rc.build_plan(d.chroms, d.genes, np.load(p["tprob"]), np.load(p["avecs"]), [d.expr, other.expr], d.H)
    assert plan.chroms == ["1", "2", "X"] and plan.S == 36 and plan.n_samples == 2
    assert plan.chains.dtype.itemsize == 32 and len(plan.chains) == 6
    assert plan.genes_per_sample == 15 and plan.expr.shape == (30, 8)
    assert list(plan.chains["n_genes"]) == [9, 5, 1] * 2
    assert list(plan.chains["n_steps"]) == [8, 5, 0] * 2
    assert list(plan.chains["gene0"]) == [0, 9, 14, 15, 24, 29]
    assert list(plan.chains["tprob0"]) == [0, 8, 13] * 2  # samples share the transition matrices
    assert list(plan.chains["state0"]) == [0, 9, 15, 16, 25, 31] and plan.n_states_out == 32
    assert plan.tprob.shape == (13, 36, 36)
    assert np.array_equal(plan.tprob[8:13], d.tprob["2"])
    assert np.array_equal(plan.init, ro.initial_logprob(8))  # same expression as the reference: bitwise
    ids = [g for c in plan.chroms for g in d.genes[c]]
    for i, g in enumerate(ids):
        a = plan.avec_index[i]
        assert (a >= 0) == (g in d.avecs) and plan.avec_index[15 + i] == a
        if a >= 0:
            assert np.array_equal(plan.avecs[a], d.avecs[g])
        assert np.array_equal(plan.expr[i], d.expr[g]) and np.array_equal(plan.expr[15 + i], other.expr[g])


def test_plan_errors_follow_the_reference(files):
    d, p = files
    expr = dict(d.expr)
    del expr[d.genes["2"][3]]
    with pytest.raises(KeyError):  # eprob[gid] of a gene the expression table lacks (gbrs_utils.py:508)
        rc.build_plan(d.chroms, d.genes, d.tprob, d.avecs, [expr], d.H)
    genes = {c: v for c, v in d.genes.items() if c != "2"}
    with pytest.raises(KeyError):  # chromosome in the transition file but not in the gene-position file (:504)
        rc.build_plan(d.chroms, genes, d.tprob, d.avecs, [d.expr], d.H)
    short = dict(d.tprob)
    short["1"] = d.tprob["1"][:5]
    with pytest.raises(IndexError):  # the reference runs off the end of tprob_c (:512)
        rc.build_plan(d.chroms, d.genes, short, d.avecs, [d.expr], d.H)
    with pytest.raises(NotImplementedError):
        rc.build_plan(d.chroms, d.genes, d.tprob, d.avecs, [d.expr], 9)


def test_there_is_no_cpu_fallback(files, monkeypatch):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    d, p = files
    monkeypatch.setenv("GBRS_DATA", p["data_dir"])
    with pytest.raises(_lib.GbrsCudaError):
        rc.reconstruct(expression_file=p["expr"], tprob_file=p["tprob"], avec_file=p["avecs"], gpos_file=p["gpos"])
    lib = _lib.load()
    assert lib.gbrs_hmm_run(1, None, 8, None, None, None, 0, None, None, None, None, None, None, None, None) == _lib.GBRS_E_CUDA
    assert lib.gbrs_hmm_emission(1, 8, None, None, None, None, 1.5, 0.12, None, None) == _lib.GBRS_E_CUDA
    assert lib.gbrs_hmm_emission(1, 9, None, None, None, None, 1.5, 0.12, None, None) == _lib.GBRS_E_ARG


def test_cli_flags_defaults_and_error_policy(tmp_path, monkeypatch, caplog):
    """Flag surface and defaults of `gbrs reconstruct` (gbrs/commands.py:153-183); exceptions are logged, exit code 0."""
    import importlib
    import logging

    from typer.testing import CliRunner

    from gbrs_b200 import commands

    runner = CliRunner()
    res = runner.invoke(commands.app, ["reconstruct", "--help"])
    assert res.exit_code == 0
    for flag in ["-e", "--expr-file", "-t", "--tprob-file", "-x", "--avec-file", "-g", "--gpos-file", "-c",
                 "--expr-threshold", "-s", "--sigma", "-o", "--outbase", "-v"]:
        assert flag in res.output, flag
    e, t = tmp_path / "e.tpm", tmp_path / "t.npz"
    e.write_text("x")
    t.write_text("x")
    seen = {}

    def fake(**kw):
        seen.update(kw)
        raise RuntimeError("boom")

    monkeypatch.setattr(importlib.import_module("gbrs_b200.reconstruct"), "reconstruct", fake)
    with caplog.at_level(logging.ERROR, logger="gbrs"):
        res = runner.invoke(commands.app, ["reconstruct", "-e", str(e), "-t", str(t)])
    assert res.exit_code == 0
    assert seen["expr_threshold"] == 1.5 and seen["sigma"] == 0.12 and seen["outbase"] is None
    assert seen["avec_file"] is None and seen["gpos_file"] is None and seen["expression_file"] == str(e)
    assert runner.invoke(commands.app, ["reconstruct", "-e", str(tmp_path / "nope"), "-t", str(t)]).exit_code != 0


def test_bulk_npz_reader_equals_numpy(tmp_path):
    """utils.read_npz_members against np.load on every kind of member the workflow meets: many small matrices, large
    3-D arrays (threaded path), string tables, Fortran order, a format-3 header; selection by list and by predicate."""
    from gbrs_b200 import utils

    rng = np.random.default_rng(0)
    small = {f"G{i:05d}": rng.random((8, 8)) for i in range(300)}
    small["odd_shape"] = rng.random((3, 5))
    small["ints"] = np.arange(7, dtype=np.int32)
    small["fortran"] = np.asfortranarray(rng.random((4, 6)))
    small["strings"] = np.array([("ENSG1", "100"), ("ENSG2", "250")])
    small["utf8_field"] = np.zeros(2, dtype=[("\u00e9", "<f8")])  # npy format 3: numpy's own reader
    p1 = tmp_path / "small.npz"
    np.savez_compressed(p1, **small)
    big = {str(c): rng.random((40, 36, 36)) for c in range(6)}
    p2 = tmp_path / "big.npz"
    np.savez_compressed(p2, **big)
    for path, ref, workers in ((p1, small, None), (p2, big, 4), (p2, big, 1)):
        got = utils.read_npz_members(str(path), None, workers=workers)
        assert list(got) == list(ref)
        for k, v in ref.items():
            assert got[k].dtype == v.dtype and got[k].shape == v.shape and np.array_equal(got[k], v), k
    assert list(utils.read_npz_members(str(p2), ["3", "1"])) == ["3", "1"]
    assert sorted(utils.read_npz_members(str(p1), lambda n: n.startswith("G0000"))) == [f"G0000{i}" for i in range(10)]
    assert utils.read_npz_members(str(p1), lambda n: False) == {}
    with pytest.raises(KeyError):
        utils.read_npz_members(str(p2), ["nope"])
