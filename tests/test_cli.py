"""CLI flag surface of `gbrs quantify` / `emase run` (reference gbrs/commands.py:108-150, emase/commands.py:315-356) and
the host-only parts of the workflow.  No GPU needed: the EM itself is stubbed where a run would start."""
import logging

import numpy as np
import pytest
from typer.testing import CliRunner

import importlib

from gbrs_b200 import cohort, commands, synth

qmod = importlib.import_module("gbrs_b200.quantify")

runner = CliRunner()


def test_quantify_help_lists_reference_flags():
    res = runner.invoke(commands.app, ["quantify", "--help"])
    assert res.exit_code == 0
    for flag in ["-i", "--alignment-file", "-g", "--group-file", "-L", "--length-file", "-G", "--genotype", "-o",
                 "--outbase", "-M", "--multiread-model", "-p", "--pseudocount", "-m", "--max-iters", "-t",
                 "--tolerance", "-a", "--report-alignment-counts", "-w", "--report-posterior", "-v"]:
        assert flag in res.output, flag


def test_run_help_lists_reference_flags():
    res = runner.invoke(commands.app, ["run", "--help"])
    assert res.exit_code == 0
    for flag in ["-i", "-g", "-L", "-o", "-M", "-p", "-l", "--read-length", "-m", "-t", "-c", "-w", "-v"]:
        assert flag in res.output, flag


def test_missing_alignment_file_is_a_usage_error(tmp_path):
    res = runner.invoke(commands.app, ["quantify", "-i", str(tmp_path / "nope.h5")])
    assert res.exit_code != 0  # typer's exists=True check, as in the reference


def test_defaults_and_error_policy(tmp_path, monkeypatch, caplog):
    """Defaults match the reference (-M 4, -p 0.0, -m 999, -t 0.0001, outbase gbrs.quantified) and exceptions are
    logged, not raised: exit code 0 (gbrs/commands.py:146-150)."""
    aln = tmp_path / "aln.npz"
    aln.write_bytes(b"x")
    seen = {}

    def fake_quantify(**kw):
        seen.update(kw)
        raise RuntimeError("boom")

    monkeypatch.setattr(commands.quantify_mod, "quantify", fake_quantify)
    with caplog.at_level(logging.ERROR, logger="gbrs"):
        res = runner.invoke(commands.app, ["quantify", "-i", str(aln)])
    assert res.exit_code == 0
    assert seen["multiread_model"] == 4 and seen["pseudocount"] == 0.0 and seen["max_iters"] == 999
    assert seen["tolerance"] == 0.0001 and seen["outbase"] == "gbrs.quantified"
    assert seen["group_file"] is None and seen["genotype_file"] is None and not seen["report_alignment_counts"]
    seen.clear()
    res = runner.invoke(commands.app, ["quantify", "-i", str(aln), "-M", "7"])
    assert res.exit_code == 0 and not seen  # rejected before the workflow starts, logged


def test_genotype_mask_matches_reference_semantics(tmp_path):
    d = synth.generate(T=40, N=200, H=8, with_genotype=True)
    apm = synth.to_apm(d)
    gt = tmp_path / "gt.tsv"
    synth.write_genotype_file(d, str(gt))
    gtmask, gtcall_g, gtcall_t = qmod.load_genotype_mask(apm, str(gt))
    assert np.array_equal(gtmask, synth.genotype_mask(d))
    assert gtcall_g[d.gname[0]] == d.genotype[0]
    assert gtcall_t[d.lname[0]] == d.genotype[int(d.gene_of[0])]
    before = apm.nnz
    apm.multiply(gtmask, axis=2)
    apm.eliminate_zeros()
    assert apm.is_pure_incidence() and apm.nnz < before


def test_genotype_mask_odd_files_follow_the_reference_loop(tmp_path):
    """Comment lines, extra columns, a gene listed twice, one- and three-letter calls, genes missing from the file
    (fully masked out), an unknown gene (KeyError): against the reference's line loop (gbrs/emase_utils.py:247-269)."""
    d = synth.generate(T=300, N=200, H=8, with_genotype=True)
    apm = synth.to_apm(d)
    gt = tmp_path / "gt.tsv"
    with open(gt, "w") as fh:
        fh.write("#Gene_ID\tDiplotype\n# another comment\n")
        for g, call in enumerate(d.genotype):
            if g % 7 != 3:
                fh.write(f"{d.gname[g]}\t{call}\tignored\n")
        fh.write(f"{d.gname[5]}\tH\n{d.gname[6]}\tABC\n")
    hid = {h: i for i, h in enumerate(apm.hname)}
    gid = {g: i for i, g in enumerate(apm.gname)}
    want, want_g, want_t = np.zeros((8, d.T)), dict.fromkeys(apm.gname), dict.fromkeys(apm.lname)
    for line in open(gt):
        if line.startswith("#"):
            continue
        g, call = line.rstrip().split("\t")[:2]
        want_g[g] = call
        for t in apm.groups[gid[g]]:
            want_t[apm.lname[t]] = call
            for c in call:
                want[hid[c], t] = 1.0
    gtmask, gtcall_g, gtcall_t = qmod.load_genotype_mask(apm, str(gt))
    assert np.array_equal(gtmask, want) and gtcall_g == want_g and gtcall_t == want_t
    assert gtcall_g[d.gname[3]] is None and gtmask[:, apm.groups[3]].sum() == 0
    with open(gt, "a") as fh:
        fh.write("NOPE\tAB\n")
    with pytest.raises(KeyError):
        qmod.load_genotype_mask(apm, str(gt))


def test_apm_npz_roundtrip_and_groups(tmp_path):
    d = synth.generate(T=30, N=150, H=4)
    apm = synth.to_apm(d)
    fn = tmp_path / "aln.emase"
    apm.save(str(fn))
    grp = tmp_path / "grp.tsv"
    synth.write_group_file(d, str(grp))
    from gbrs_b200 import AlignmentPropertyMatrix

    back = AlignmentPropertyMatrix(h5file=str(fn), grpfile=str(grp))
    assert back.shape == apm.shape and back.hname == list(d.hname) and list(back.lname) == d.lname
    assert np.array_equal(back.count, d.count) and back.num_groups == len(d.gname)
    for a, b in zip(apm.data, back.data):
        assert np.array_equal(a.indptr, b.indptr) and np.array_equal(a.indices, b.indices)
    assert back.groups == d.groups()


def test_cohort_round_robin():
    assert cohort.my_share(10, 0, 4) == [0, 4, 8] and cohort.my_share(10, 3, 4) == [3, 7]
    assert sorted(sum((cohort.my_share(96, r, 8) for r in range(8)), [])) == list(range(96))
    with pytest.raises(ValueError):
        cohort.my_share(4, 2, 2)
