"""`quantify` end to end on the GPU against the text tables written by the reference's own quantify()."""
import os

import numpy as np
import pytest

from gbrs_b200 import synth
from gbrs_b200.quantify import quantify
from tests import helpers as hp
from tests.test_pack import make_apm

pytestmark = pytest.mark.gpu


def read_table(path):
    with open(path) as fh:
        header = fh.readline().rstrip("\n").split("\t")
        names, rows, notes = [], [], []
        ncol = len(header) - 1 - (header[-1] == "notes")
        for line in fh:
            it = line.rstrip("\n").split("\t")
            names.append(it[0])
            rows.append([float(x) for x in it[1:1 + ncol]])
            notes.append(it[1 + ncol] if header[-1] == "notes" else None)
    return header, names, np.array(rows), notes


@pytest.mark.parametrize("case,kind,genotype", [("quantify_multiway", "multiway", False),
                                                ("quantify_diploid", "diploid", True),
                                                ("quantify_multiway_m2", "multiway", False)])
def test_quantify_tables_match_reference(case, kind, genotype, tmp_path, capsys):
    gdir = os.path.join(hp.GOLDEN, case)
    z = np.load(os.path.join(gdir, "input.npz"))
    d = synth.generate(T=int(z["T"]), N=int(z["N"]), H=int(z["H"]), with_genotype=True)
    if not (np.array_equal(d.pair_class, z["pair_class"]) and np.array_equal(d.pair_mask, z["pair_mask"])):
        d.pair_class, d.pair_locus = z["pair_class"].astype(np.int64), z["pair_locus"].astype(np.int64)
        d.pair_mask, d.count = z["pair_mask"], z["count"]
    apm = make_apm(d)
    aln = os.path.join(str(tmp_path), "aln.emase")
    apm.save(aln)
    outbase = os.path.join(str(tmp_path), "out")
    quantify(alignment_file=aln, group_file=os.path.join(gdir, "grp.tsv"), length_file=os.path.join(gdir, "len.tsv"),
             genotype_file=os.path.join(gdir, "gt.tsv") if genotype else None, outbase=outbase,
             multiread_model=int(z["model"]), pseudocount=float(z["pseudocount"]), max_iters=999, tolerance=1e-4,
             report_alignment_counts=True)
    # same iteration table length as the reference
    ref_iters = sum("/ 1000000" in line for line in open(os.path.join(gdir, "stdout.txt")))
    assert sum("/ 1000000" in line for line in capsys.readouterr().out.splitlines()) == ref_iters
    suffixes = ["isoforms.tpm", "isoforms.expected_read_counts", "genes.tpm", "genes.expected_read_counts",
                "isoforms.alignment_counts", "genes.alignment_counts"]
    for sfx in suffixes:
        ref = read_table(os.path.join(gdir, f"out.{kind}.{sfx}"))
        got = read_table(f"{outbase}.{kind}.{sfx}")
        assert got[0] == ref[0], sfx
        assert got[1] == ref[1], sfx
        assert got[3] == ref[3], sfx
        if sfx.endswith("alignment_counts"):
            assert np.array_equal(got[2], ref[2]), sfx
        else:
            scale = np.abs(ref[2]).max()
            assert np.abs(got[2] - ref[2]).max() <= 1e-9 * scale, sfx
