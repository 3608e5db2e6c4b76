"""`compress` (equivalence classes of reads): host re-layout on the CPU, the GPU grouping through the C-ABI against the
reference-generated golden vectors and against the oracle on seeded inputs."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

from gbrs_b200 import compress as cz
from gbrs_b200.apm import AlignmentPropertyMatrix
from oracle import compress_oracle as co
from oracle.make_golden_compress import make_reads
from tests.test_compress_oracle import CASES, load_case, same_pattern


def write_inputs(tmp_path, files, T, H):
    hname = [chr(ord("A") + h) for h in range(H)]
    lname = [f"T{t:04d}" for t in range(T)]
    paths = []
    for i, (mats, count) in enumerate(files):
        p = os.path.join(str(tmp_path), f"in{i}.npz")
        AlignmentPropertyMatrix.from_csc(mats, hname, lname, count=count).save(h5file=p)
        paths.append(p)
    return paths


def test_read_rows_carry_the_reference_key():
    """The pair-word row of a read holds exactly the information of the reference's string key (per haplotype the
    sorted locus ids, emase_utils.py:62-71), so equal rows <=> equal keys."""
    mats, _ = make_reads(seed=11, T=30, H=5, n_classes=40, n_reads=300, with_count=False, empty_reads=4)
    rowptr, words = cz.read_rows(mats, 30, 5)
    keys = co.read_keys(mats)
    assert rowptr[-1] == len(words) and len(rowptr) == 305
    for r, key in enumerate(keys):
        w = words[rowptr[r]:rowptr[r + 1]].astype(np.int64)
        loci, mask = w & 0xFFFFFF, w >> 24
        assert np.all(np.diff(loci) > 0) and np.all(mask > 0)
        for h in range(5):
            assert tuple(loci[(mask >> h) & 1 == 1]) == key[h]
    # rows equal <=> keys equal
    rows = [tuple(words[rowptr[r]:rowptr[r + 1]]) for r in range(304)]
    assert len(set(rows)) == len(set(keys))
    # representative rows expand back into the per-haplotype matrices
    first = np.array([rows.index(x) for x in dict.fromkeys(rows)], dtype=np.uint32)
    back = cz.class_matrices(rowptr, words, first, 30, 5)
    want, _ = co.compress([(mats, None)])
    for h in range(5):
        assert same_pattern(back[h], want[h])


def test_explicit_zeros_are_not_alignments():
    m = sp.csc_matrix(np.array([[1.0, 0.0], [1.0, 1.0]]))
    m.data[0] = 0.0  # stored zero at (0, 0)
    rowptr, words = cz.read_rows([m], 2, 1)
    assert list(rowptr) == [0, 0, 2] and [int(w) & 0xFFFFFF for w in words] == [0, 1]


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_gpu_compress_matches_reference_golden(name, tmp_path):
    T, H, files, want, want_count = load_case(name)
    out = os.path.join(str(tmp_path), "out.npz")
    cz.compress(write_inputs(tmp_path, files, T, H), out)
    res = AlignmentPropertyMatrix(h5file=out)
    assert res.shape == (T, H, len(want_count))
    assert np.array_equal(res.count, want_count)  # integer-valued sums: bit-exact
    for h in range(H):
        assert same_pattern(res.data[h], want[h])  # same classes in the same (first-appearance) order


@pytest.mark.gpu
def test_gpu_equivalence_classes_match_oracle_at_size():
    """200k reads drawn from 20k patterns in two files: class order, membership and counts against the oracle."""
    files = [make_reads(seed=21, T=3000, H=8, n_classes=20000, n_reads=120000, with_count=True, empty_reads=7),
             make_reads(seed=22, T=3000, H=8, n_classes=20000, n_reads=80000, with_count=False)]
    rps, ws, cs, base = [], [], [], 0
    for mats, count in files:
        rp, w = cz.read_rows(mats, 3000, 8)
        rps.append(rp[1:] + base if rps else rp)
        base += int(rp[-1])
        ws.append(w)
        cs.append(np.ones(mats[0].shape[0]) if count is None else count)
    rowptr, words, count = np.concatenate(rps), np.concatenate(ws), np.concatenate(cs)
    cls, first, ccount = cz.equivalence_classes(rowptr, words, count)
    want_mats, want_count = co.compress(files)
    assert len(first) == len(want_count) and np.array_equal(ccount, want_count)
    got = cz.class_matrices(rowptr, words, first, 3000, 8)
    for h in range(8):
        assert same_pattern(got[h], want_mats[h])
    # membership: every read sits in the class of the first read with its row; class ids ascend with first appearance
    assert np.all(np.diff(first.astype(np.int64)) > 0) and np.array_equal(cls[first], np.arange(len(first)))
    rows = {}
    for r in range(0, len(cls), 97):
        key = tuple(words[rowptr[r]:rowptr[r + 1]])
        assert tuple(words[rowptr[first[cls[r]]]:rowptr[first[cls[r]] + 1]]) == key
        rows.setdefault(key, cls[r])
        assert rows[key] == cls[r]
    # identical input, identical answer (the atomic count sums are integer valued)
    cls2, first2, ccount2 = cz.equivalence_classes(rowptr, words, count)
    assert np.array_equal(cls, cls2) and np.array_equal(first, first2) and np.array_equal(ccount, ccount2)


@pytest.mark.gpu
def test_gpu_compress_edge_cases(tmp_path):
    # only empty reads; a single read; all reads identical
    for mats, count, n_ec in [
        ([sp.csc_matrix((5, 4)) for _ in range(2)], None, 1),
        ([sp.csc_matrix(np.array([[0, 1.0, 0, 1.0]])) for _ in range(2)], np.array([3.0]), 1),
        ([sp.csc_matrix(np.tile(np.array([[1.0, 0, 0, 1.0]]), (6, 1))) for _ in range(2)], None, 1),
    ]:
        rowptr, words = cz.read_rows(mats, 4, 2)
        cls, first, ccount = cz.equivalence_classes(rowptr, words, count)
        want_mats, want_count = co.compress([(mats, count)])
        assert len(first) == n_ec == len(want_count) and np.array_equal(ccount, want_count) and first[0] == 0
        assert np.all(cls == 0)


@pytest.mark.gpu
def test_cli_compress(tmp_path):
    from typer.testing import CliRunner

    from gbrs_b200.commands import app

    T, H, files, want, want_count = load_case("compress_two_files_counts")
    paths = write_inputs(tmp_path, files, T, H)
    out = os.path.join(str(tmp_path), "merged.npz")
    res = CliRunner().invoke(app, ["compress", "-i", ",".join(paths), "-o", out])
    assert res.exit_code == 0, res.output
    got = AlignmentPropertyMatrix(h5file=out)
    assert np.array_equal(got.count, want_count) and all(same_pattern(got.data[h], want[h]) for h in range(H))
    # a missing input is logged, not raised (commands.py:101-105)
    res = CliRunner().invoke(app, ["compress", "-i", os.path.join(str(tmp_path), "nope.h5"), "-o", out + "2"])
    assert res.exit_code == 0 and not os.path.exists(out + "2")


def _dict_grouping(rowptr, words, count=None, device=None):
    """Test-only stand-in for the device step (`equivalence_classes`): the same contract from a python dict."""
    n = len(rowptr) - 1
    seen, cls, first, ccount = {}, np.zeros(n, np.uint32), [], []
    for r in range(n):
        key = words[rowptr[r]:rowptr[r + 1]].tobytes()
        c = seen.setdefault(key, len(seen))
        if c == len(first):
            first.append(r)
            ccount.append(0.0)
        cls[r] = c
        ccount[c] += 1.0 if count is None else float(count[r])
    return cls, np.array(first, dtype=np.uint32), np.array(ccount)


@pytest.mark.parametrize("name", CASES)
def test_compress_workflow_on_cpu_with_the_device_step_replaced(name, tmp_path, monkeypatch):
    """Everything of `compress` around the GPU grouping -- reading several files in order, read rows, counts,
    representative rows back into matrices, the output file, the CLI -- against the reference-written goldens."""
    from typer.testing import CliRunner

    from gbrs_b200.commands import app

    monkeypatch.setattr(cz, "equivalence_classes", _dict_grouping)
    T, H, files, want, want_count = load_case(name)
    paths = write_inputs(tmp_path, files, T, H)
    out = os.path.join(str(tmp_path), "out.npz")
    cz.compress(paths, out)
    res = AlignmentPropertyMatrix(h5file=out)
    assert res.shape == (T, H, len(want_count)) and np.array_equal(res.count, want_count)
    assert all(same_pattern(res.data[h], want[h]) for h in range(H))
    out2 = os.path.join(str(tmp_path), "cli.npz")
    args = ["compress", "-o", out2]
    for p in paths:
        args += ["-i", p]
    assert CliRunner().invoke(app, args).exit_code == 0
    got = AlignmentPropertyMatrix(h5file=out2)
    assert np.array_equal(got.count, want_count) and all(same_pattern(got.data[h], want[h]) for h in range(H))


@pytest.mark.parametrize("name", CASES)
def test_emulated_ec_kernels_match_reference_golden(name, tmp_path, monkeypatch):
    """The GPU grouping code itself (ec_kernels.cu: hashing, exact comparison of sorted neighbours, first-appearance
    numbering, class counts, launcher, C ABI) on the host SIMT shim -- the CUB sort / scan calls served by a stable sort
    and a loop -- behind the unchanged `compress` workflow, against the goldens written by the reference."""
    import shutil

    if shutil.which("g++") is None:
        pytest.skip("g++ is needed to build the SIMT emulation")
    from tests import simt_em

    monkeypatch.setattr(cz, "equivalence_classes", simt_em.emulated_equivalence_classes)
    T, H, files, want, want_count = load_case(name)
    out = os.path.join(str(tmp_path), "out.npz")
    cz.compress(write_inputs(tmp_path, files, T, H), out)
    res = AlignmentPropertyMatrix(h5file=out)
    assert res.shape == (T, H, len(want_count)) and np.array_equal(res.count, want_count)
    assert all(same_pattern(res.data[h], want[h]) for h in range(H))


def test_emulated_ec_kernels_match_oracle_and_edge_cases():
    import shutil

    if shutil.which("g++") is None:
        pytest.skip("g++ is needed to build the SIMT emulation")
    from tests import simt_em

    files = [make_reads(seed=21, T=300, H=8, n_classes=900, n_reads=5000, with_count=True, empty_reads=7),
             make_reads(seed=22, T=300, H=8, n_classes=900, n_reads=3000, with_count=False)]
    rps, ws, cs, base = [], [], [], 0
    for mats, count in files:
        rp, w = cz.read_rows(mats, 300, 8)
        rps.append(rp[1:] + base if rps else rp)
        base += int(rp[-1])
        ws.append(w)
        cs.append(np.ones(mats[0].shape[0]) if count is None else count)
    rowptr, words, count = np.concatenate(rps), np.concatenate(ws), np.concatenate(cs)
    cls, first, ccount = simt_em.emulated_equivalence_classes(rowptr, words, count)
    want_mats, want_count = co.compress(files)
    assert len(first) == len(want_count) and np.array_equal(ccount, want_count)
    got = cz.class_matrices(rowptr, words, first, 300, 8)
    assert all(same_pattern(got[h], want_mats[h]) for h in range(8))
    assert np.all(np.diff(first.astype(np.int64)) > 0) and np.array_equal(cls[first], np.arange(len(first)))
    for mats, cnt, n_ec in [([sp.csc_matrix((5, 4)) for _ in range(2)], None, 1),
                            ([sp.csc_matrix(np.array([[0, 1.0, 0, 1.0]])) for _ in range(2)], np.array([3.0]), 1),
                            ([sp.csc_matrix(np.tile(np.array([[1.0, 0, 0, 1.0]]), (6, 1))) for _ in range(2)], None, 1)]:
        rp, w = cz.read_rows(mats, 4, 2)
        c, f, cc = simt_em.emulated_equivalence_classes(rp, w, cnt)
        _, wc = co.compress([(mats, cnt)])
        assert len(f) == n_ec == len(wc) and np.array_equal(cc, wc) and f[0] == 0 and np.all(c == 0)
