"""`stencil` (genotype calls applied to an alignment incidence file) against a golden vector written by the unmodified
reference (oracle/make_golden_stencil.py).  Host-only step: runs without a GPU."""
import os

import numpy as np
import scipy.sparse as sp

from gbrs_b200 import synth
from gbrs_b200.apm import AlignmentPropertyMatrix
from gbrs_b200.stencil import stencil
from tests.test_compress_oracle import same_pattern
from tests.test_pack import make_apm

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "stencil_small.npz")


def setup_case(tmp_path):
    g = np.load(GOLDEN)
    T, H, N = int(g["T"]), int(g["H"]), int(g["N"])
    d = synth.generate(T=T, N=N, H=H, with_genotype=True, sample_index=int(g["sample_index"]))
    assert np.array_equal(d.pair_mask, g["pair_mask"]) and np.array_equal(d.pair_locus, g["pair_locus"])  # same input
    aln, grp, gt = (os.path.join(str(tmp_path), x) for x in ("aln.npz", "grp.tsv", "gt.tsv"))
    make_apm(d).save(h5file=aln)
    open(grp, "w").write(str(g["group_tsv"]))
    open(gt, "w").write(str(g["genotype_tsv"]))
    want = [sp.csc_matrix((np.ones(len(g[f"out_h{h}_indices"])), g[f"out_h{h}_indices"], g[f"out_h{h}_indptr"]),
                          shape=(N, T)) for h in range(H)]
    return d, aln, grp, gt, want, g["count"]


def test_stencil_matches_reference(tmp_path):
    d, aln, grp, gt, want, want_count = setup_case(tmp_path)
    out = os.path.join(str(tmp_path), "stenciled.npz")
    stencil(alignment_file=aln, genotype_file=gt, group_file=grp, output_file=out)
    res = AlignmentPropertyMatrix(h5file=out)
    assert res.shape == (d.T, d.H, d.N) and np.array_equal(res.count, want_count)
    for h in range(d.H):
        assert same_pattern(res.data[h], want[h])
    assert sum(m.nnz for m in res.data) < d.nnz  # something was masked out


def test_stencil_cli_and_default_output_name(tmp_path, monkeypatch):
    from typer.testing import CliRunner

    from gbrs_b200.commands import app

    d, aln, grp, gt, want, _ = setup_case(tmp_path)
    monkeypatch.chdir(tmp_path)
    res = CliRunner().invoke(app, ["stencil", "-i", aln, "-G", gt, "-g", grp])
    assert res.exit_code == 0, res.output
    got = AlignmentPropertyMatrix(h5file=os.path.join(str(tmp_path), "gbrs.stenciled.aln.npz"))
    assert all(same_pattern(got.data[h], want[h]) for h in range(d.H))


def test_stencil_without_groups_is_keyed_by_locus(tmp_path, monkeypatch):
    d, aln, _, _, _, _ = setup_case(tmp_path)
    monkeypatch.setenv("GBRS_DATA", str(tmp_path / "nowhere"))
    gt = os.path.join(str(tmp_path), "gt_loci.tsv")
    with open(gt, "w") as fh:
        fh.write("#Locus\tDiplotype\n")
        for t in range(0, d.T, 2):
            fh.write(f"{d.lname[t]}\tAC\n")
    out = os.path.join(str(tmp_path), "o.npz")
    stencil(alignment_file=aln, genotype_file=gt, output_file=out)
    res = AlignmentPropertyMatrix(h5file=out)
    for h in range(d.H):
        per_locus = np.diff(res.data[h].indptr)
        assert np.all(per_locus[1::2] == 0)
        assert (h in (0, 2)) or np.all(per_locus == 0)
    keep = ((d.pair_mask & 0b101) != 0) & (d.pair_locus % 2 == 0)
    assert sum(m.nnz for m in res.data) == int(sum(bin(int(m) & 0b101).count("1") for m in d.pair_mask[keep]))
