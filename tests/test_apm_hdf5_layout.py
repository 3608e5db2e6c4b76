"""The PyTables (HDF5) branch of `AlignmentPropertyMatrix` load / save -- the format real EMASE files are in.  PyTables
and libhdf5 are not installed in the build image, so the branch is exercised against a stand-in `tables` module that
implements exactly the calls the reference makes (Sparse3DMatrix.py:42-50, :68-102, :400-444;
AlignmentPropertyMatrix.py:70-83, :478-525) with PyTables' observable behaviour: AttributeError for a missing attribute,
str arrays stored as bytes, `in` on paths like `//count`.  What this pins: node names, attribute names, dtypes and the
order of calls -- i.e. that a file written by the reference is read and a file written here is readable by the
reference.  What it cannot pin: the HDF5 byte format itself (PyTables' job)."""
import os
import pickle
import re
import sys
import types

import numpy as np
import pytest
import scipy.sparse as sp

from gbrs_b200 import synth
from gbrs_b200.apm import _HDF5_MAGIC, AlignmentPropertyMatrix as APM


def _norm(path):
    return re.sub(r"/+", "/", "/" + path.strip()) if path.strip("/") else "/"


class _Node:
    def __init__(self, fh, path):
        self._fh, self._path = fh, _norm(path)

    def read(self):
        return self._fh.store["nodes"][self._path]


class _File:
    """An "HDF5" file = the HDF5 signature followed by a pickle of {attrs, nodes}."""

    def __init__(self, filename, mode="r", title=None, **kw):
        self.filename, self.mode = filename, mode
        if mode == "w":
            self.store = {"attrs": {}, "nodes": {}, "title": title}
        else:
            with open(filename, "rb") as fh:
                assert fh.read(8) == _HDF5_MAGIC
                self.store = pickle.load(fh)
        self.root = _Node(self, "/")
        self.calls = []

    def _p(self, where, name=None):
        base = where._path if isinstance(where, _Node) else str(where)
        return _norm(base + ("/" + name if name else ""))

    def get_node_attr(self, where, name):
        try:
            return self.store["attrs"][self._p(where)][name]
        except KeyError:
            raise AttributeError(name) from None

    def set_node_attr(self, where, name, value):
        self.store["attrs"].setdefault(self._p(where), {})[name] = value

    def get_node(self, where, name=None):
        p = self._p(where, name)
        if p not in self.store["nodes"] and p not in self.store["attrs"]:
            raise LookupError(p)  # tables.NoSuchNodeError
        return _Node(self, p)

    def __contains__(self, path):
        p = _norm(path)
        return p in self.store["nodes"] or p in self.store["attrs"]

    def create_group(self, where, name, title=None):
        p = self._p(where, name)
        self.store["attrs"].setdefault(p, {})
        return _Node(self, p)

    def create_carray(self, where, name, obj=None, title=None, filters=None):
        arr = np.asarray(obj)
        if arr.dtype.kind == "U":
            arr = arr.astype("S")
        self.store["nodes"][self._p(where, name)] = arr
        return _Node(self, self._p(where, name))

    def flush(self):
        if self.mode == "w":
            with open(self.filename, "wb") as fh:
                fh.write(_HDF5_MAGIC)
                pickle.dump(self.store, fh)

    def close(self):
        self.flush()


@pytest.fixture()
def fake_tables(monkeypatch):
    m = types.ModuleType("tables")
    m.open_file = lambda filename, mode="r", title=None, **kw: _File(filename, mode, title)
    m.Filters = lambda **kw: ("filters", kw)
    monkeypatch.setitem(sys.modules, "tables", m)
    return m


def same(a, b):
    assert a.shape == b.shape and a.hname == list(b.hname) and list(a.lname) == list(b.lname)
    assert (a.count is None) == (b.count is None) and (a.count is None or np.array_equal(a.count, b.count))
    for x, y in zip(a.data, b.data):
        x, y = sp.csc_matrix(x), sp.csc_matrix(y)
        x.sort_indices(), y.sort_indices()
        assert np.array_equal(x.indptr, y.indptr) and np.array_equal(x.indices, y.indices) and np.array_equal(x.data, y.data)


def test_hdf5_branch_round_trip(tmp_path, fake_tables):
    d = synth.generate(T=40, N=300, H=8)
    apm = synth.to_apm(d)
    fn = str(tmp_path / "aln.h5")
    apm.save(h5file=fn, title="t")  # `tables` importable and not an .npz name: the PyTables branch
    with open(fn, "rb") as fh:
        assert fh.read(8) == _HDF5_MAGIC
        store = pickle.load(fh)
    # layout of AlignmentPropertyMatrix.save / Sparse3DMatrix.save in the reference
    assert store["attrs"]["/"]["mtype"] == "csc_matrix" and store["attrs"]["/"]["incidence_only"] is True
    assert tuple(store["attrs"]["/"]["shape"]) == (40, 8, 300) and list(store["attrs"]["/"]["hname"]) == list(d.hname)
    assert set(store["nodes"]) == {f"/h{h}/{k}" for h in range(8) for k in ("indptr", "indices")} | {"/count", "/lname"}
    assert store["nodes"]["/h0/indptr"].dtype == np.uint32 and store["nodes"]["/h3/indices"].dtype == np.uint32
    assert store["nodes"]["/lname"].dtype.kind == "S" and store["nodes"]["/count"].dtype == np.float64
    back = APM(h5file=fn)  # signature sniffing -> PyTables branch
    same(back, apm)
    assert back.finalized and back.lid[d.lname[7]] == 7 and back.is_pure_incidence()
    # values, read names, no counts, shallow
    apm2 = synth.to_apm(d)
    apm2.count = None
    apm2.rname = np.array([f"r{i}".encode() for i in range(d.N)])
    apm2.data[2] = apm2.data[2] * 2.5
    fn2 = str(tmp_path / "values.h5")
    apm2.save(h5file=fn2, incidence_only=False)
    back2 = APM(h5file=fn2)
    same(back2, apm2)
    assert back2.count is None and list(back2.rname[:2]) == [b"r0", b"r1"] and back2.rid[b"r5"] == 5
    shallow = APM(h5file=fn, shallow=True)
    assert shallow.hname is None and shallow.lname is None and np.array_equal(shallow.count, apm.count)


def test_hdf5_branch_reads_the_legacy_coo_layout(tmp_path, fake_tables):
    """Files without `mtype` / `incidence_only` attributes hold COO components (Sparse3DMatrix.py:69-78, :96-102)."""
    d = synth.generate(T=12, N=60, H=2)
    mats = synth.to_csc_list(d)
    fn = str(tmp_path / "legacy.h5")
    fh = _File(fn, "w")
    fh.set_node_attr(fh.root, "shape", (12, 2, 60))
    fh.set_node_attr(fh.root, "hname", ["A", "B"])
    for h, m in enumerate(mats):
        g = fh.create_group(fh.root, f"h{h}")
        c = m.tocoo()
        fh.create_carray(g, "coor", obj=np.vstack((c.row, c.col)).astype(np.uint32))
        fh.create_carray(g, "data", obj=c.data)
    fh.create_carray(fh.root, "lname", obj=np.array(d.lname))
    fh.close()
    back = APM(h5file=fn)
    assert back.count is None
    for h in range(2):
        assert (sp.csc_matrix(back.data[h]) != mats[h]).nnz == 0


def test_without_pytables_an_hdf5_file_is_refused_with_a_pointer_to_the_converter(tmp_path, monkeypatch):
    monkeypatch.setitem(sys.modules, "tables", None)  # import tables -> ImportError
    fn = tmp_path / "real.h5"
    fn.write_bytes(_HDF5_MAGIC + b"\0" * 64)
    with pytest.raises(RuntimeError, match="gbrs_b200.convert"):
        APM(h5file=str(fn))


@pytest.mark.skipif(not os.path.isdir("/root/reference/src/gbrs/emase"), reason="needs the reference sources")
def test_hdf5_calls_are_interchangeable_with_the_reference(tmp_path, monkeypatch):
    """Both directions through ONE stand-in `tables`: the reference writes / we read, we write / the reference reads."""
    from oracle import ref_harness as rh

    # load_reference() registers the oracle's pickle-backed `tables` for good; keep that out of the other tests
    monkeypatch.setitem(sys.modules, "tables", rh._fake_tables_module())
    ref = rh.load_reference()  # imports the reference with a stand-in `tables` (its modules keep their own binding)
    d = synth.generate(T=30, N=200, H=8, sample_index=3)
    theirs = rh.build_reference_apm(d)
    f1 = str(tmp_path / "by_reference.h5")
    theirs.save(h5file=f1)
    monkeypatch.setitem(sys.modules, "tables", ref.apm_mod.tables)  # the module the reference itself is bound to
    ours = APM()
    ours._load_hdf5(f1, "/", "/", False, float)  # the stand-in's files carry no HDF5 signature: call the branch directly
    ours.num_loci, ours.num_haplotypes, ours.num_reads = ours.shape
    same(ours, synth.to_apm(d))
    f2 = str(tmp_path / "by_us.h5")
    mine = synth.to_apm(d)
    mine._save_hdf5(ref.apm_mod.tables, f2, None, "uint32", float, True, "zlib", False)
    back = ref.APM(h5file=f2)
    assert tuple(back.shape) == (30, 8, 200) and list(back.hname) == list(d.hname) and list(back.lname) == list(d.lname)
    assert np.array_equal(back.count, d.count)
    for h in range(8):
        assert (sp.csc_matrix(back.data[h]) != mine.data[h]).nnz == 0


def test_converter_both_directions(tmp_path, fake_tables, monkeypatch):
    from gbrs_b200 import convert

    d = synth.generate(T=25, N=120, H=4)
    apm = synth.to_apm(d)
    h5, npz, h5b = (str(tmp_path / n) for n in ("a.h5", "a.npz", "b.h5"))
    apm.save(h5file=h5)
    assert convert.main([h5, npz]) == 0 and convert.main([npz, h5b]) == 0
    with open(npz, "rb") as fh:
        assert fh.read(2) == b"PK"
    same(APM(h5file=npz), apm)
    same(APM(h5file=h5b), apm)
    assert convert.main([h5]) == 2  # usage
    monkeypatch.setitem(sys.modules, "tables", None)
    with pytest.raises(RuntimeError, match="PyTables"):
        convert.convert(npz, str(tmp_path / "c.h5"))
