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


def test_cli_quantify_and_run_end_to_end(tmp_path):
    """The typer commands drive the GPU path and write the reference's file set."""
    from typer.testing import CliRunner

    from gbrs_b200 import commands

    gdir = os.path.join(hp.GOLDEN, "quantify_multiway")
    z = np.load(os.path.join(gdir, "input.npz"))
    d = synth.generate(T=int(z["T"]), N=int(z["N"]), H=int(z["H"]), with_genotype=True)
    aln = os.path.join(str(tmp_path), "aln.emase")
    make_apm(d).save(aln)
    outbase = os.path.join(str(tmp_path), "cli")
    res = CliRunner().invoke(commands.app, ["quantify", "-i", aln, "-g", os.path.join(gdir, "grp.tsv"), "-L",
                                            os.path.join(gdir, "len.tsv"), "-o", outbase, "-M", "4", "-a"])
    assert res.exit_code == 0, res.output
    for sfx in ["isoforms.tpm", "isoforms.expected_read_counts", "genes.tpm", "genes.expected_read_counts",
                "isoforms.alignment_counts", "genes.alignment_counts"]:
        ref = read_table(os.path.join(gdir, f"out.multiway.{sfx}"))
        got = read_table(f"{outbase}.multiway.{sfx}")
        assert got[0] == ref[0] and got[1] == ref[1]
        assert np.abs(got[2] - ref[2]).max() <= 1e-9 * np.abs(ref[2]).max()
    # `emase run`: same EM, explicit read length, -c for alignment counts, no .multiway suffix
    out2 = os.path.join(str(tmp_path), "em")
    res = CliRunner().invoke(commands.app, ["run", "-i", aln, "-g", os.path.join(gdir, "grp.tsv"), "-L",
                                            os.path.join(gdir, "len.tsv"), "-o", out2, "-l", "100", "-c"])
    assert res.exit_code == 0, res.output
    a = read_table(f"{out2}.isoforms.expected_read_counts")
    b = read_table(f"{outbase}.multiway.isoforms.expected_read_counts")
    assert np.array_equal(a[2], b[2])
    assert os.path.exists(f"{out2}.genes.alignment_counts")


def test_cohort_mode_equals_individual_runs(tmp_path):
    from gbrs_b200 import cohort
    from gbrs_b200.emfactory import EMfactory

    samples = [0, 1, 2]
    lenfile = os.path.join(str(tmp_path), "len.tsv")
    base = synth.generate(T=300, N=4000, H=8, sample_index=0)
    synth.write_length_file(base, lenfile)

    def load(i):
        return make_apm(synth.generate(T=300, N=4000, H=8, sample_index=i))

    res = cohort.quantify_cohort(samples, load, model=4, lenfile=lenfile)
    assert sorted(res) == samples
    for i in samples:
        em = EMfactory(load(i))
        em.prepare(lenfile=lenfile)
        em.run(model=4, tol=1e-4, verbose=False)
        assert res[i]["iters"] == em.num_iters and np.array_equal(res[i]["theta"], em.allelic_expression)
    # different samples give different answers (they share loci and lengths only)
    assert not np.array_equal(res[0]["theta"], res[1]["theta"])
    assert [r["iters"] for r in cohort.gather_results(res, 3)] == [res[i]["iters"] for i in samples]


def test_diploid_fast_mask_path_equals_host_masking(tmp_path):
    """`quantify -G` applies the genotype restriction while packing (one byte per locus); with `-w` it masks the host
    matrices like the reference does (the exported pattern must be the restricted one).  Same tables either way."""
    gdir = os.path.join(hp.GOLDEN, "quantify_diploid")
    z = np.load(os.path.join(gdir, "input.npz"))
    d = synth.generate(T=int(z["T"]), N=int(z["N"]), H=int(z["H"]), with_genotype=True)
    aln = os.path.join(str(tmp_path), "aln.emase")
    make_apm(d).save(aln)
    outs = []
    for posterior in (False, True):
        outbase = os.path.join(str(tmp_path), f"o{int(posterior)}")
        quantify(alignment_file=aln, group_file=os.path.join(gdir, "grp.tsv"), length_file=os.path.join(gdir, "len.tsv"),
                 genotype_file=os.path.join(gdir, "gt.tsv"), outbase=outbase, multiread_model=4,
                 report_posterior=posterior)
        outs.append(read_table(f"{outbase}.diploid.isoforms.expected_read_counts"))
    assert outs[0][1] == outs[1][1] and np.array_equal(outs[0][2], outs[1][2])
    from gbrs_b200 import AlignmentPropertyMatrix

    exported = AlignmentPropertyMatrix(h5file=os.path.join(str(tmp_path), "o1.diploid.posterior.h5"))
    assert exported.nnz < make_apm(d).nnz  # the exported pattern is the restricted one
