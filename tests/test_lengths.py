"""Length-file handling of EMfactory.prepare (reference EMfactory.py:60-94): the bulk parser and the reference-style
line loop give the same table; malformed files raise what the loop raises."""
import os

import numpy as np
import pytest

from gbrs_b200 import synth
from gbrs_b200.emfactory import EMfactory
from tests.test_pack import make_apm


def factory(d):
    em = EMfactory.__new__(EMfactory)
    em.probability = make_apm(d)
    return em


@pytest.mark.parametrize("H", [8, 3, 1])
def test_bulk_parser_equals_line_loop(H, tmp_path):
    d = synth.generate(T=300, N=2000, H=H)
    lenfile = os.path.join(str(tmp_path), "len.tsv")
    synth.write_length_file(d, lenfile)
    with open(lenfile, "a") as fh:  # a duplicate key (later line wins) and a length below the read length (-> 1.0)
        name = d.lname[5] + ("_" + d.hname[0] if H > 1 else "")
        fh.write(f"{name}\t777\n{d.lname[7] + ('_' + d.hname[H - 1] if H > 1 else '')}\t40\n")
    em = factory(d)
    bulk = em._read_lengths_bulk(lenfile, 100)
    loop = em._read_lengths_loop(lenfile, 100)
    assert bulk is not None
    assert np.array_equal(bulk, loop) and bulk[5, 0] == 678.0 and bulk[7, H - 1] == 1.0
    assert np.array_equal(em._read_lengths(lenfile, 100), loop.T)


def test_malformed_length_files_raise_like_the_reference(tmp_path):
    d = synth.generate(T=20, N=100, H=2)
    em = factory(d)
    ok = os.path.join(str(tmp_path), "ok.tsv")
    synth.write_length_file(d, ok)
    lines = open(ok).read().splitlines()
    cases = {"missing": lines[:-1], "unknown_locus": lines + ["NOPE_A\t500"], "bad_key": lines + ["T_0_A\t500"],
             "not_a_number": lines[:-1] + [lines[-1].split("\t")[0] + "\tabc"]}
    want = {"missing": RuntimeError, "unknown_locus": KeyError, "bad_key": ValueError, "not_a_number": ValueError}
    for name, content in cases.items():
        f = os.path.join(str(tmp_path), name + ".tsv")
        open(f, "w").write("\n".join(content) + "\n")
        with pytest.raises(want[name]):
            em._read_lengths(f, 100)
    # things python's float() takes but the native parser leaves to the loop: same table either way
    f = os.path.join(str(tmp_path), "odd_numbers.tsv")
    open(f, "w").write("\n".join(lines[:-2] + [lines[-2].split("\t")[0] + "\t1_000", lines[-1].split("\t")[0] + "\t 5e2 "]) + "\n")
    assert em._read_lengths_bulk(f, 100) is None
    tl = em._read_lengths(f, 100)
    assert tl[1, 19] == 401.0 and tl[0, 19] == 901.0
    # windows line ends and extra columns are fine for both
    f = os.path.join(str(tmp_path), "crlf.tsv")
    open(f, "w").write("\r\n".join(x + "\textra" for x in lines) + "\r\n")
    assert np.array_equal(em._read_lengths_bulk(f, 100), em._read_lengths_loop(f, 100))
