"""TEST INFRASTRUCTURE ONLY -- runs the *unmodified* reference (churchill-lab/gbrs, mounted read-only at
/root/reference) inside this container so that golden vectors can be generated and the oracle
restatement (oracle/em_oracle.py) can be pinned against it.

Nothing in the product package imports this file.  It only works where /root/reference exists (the
build container); the GPU box never sees it -- tests there use the committed fixtures in
tests/golden/ produced by oracle/make_golden.py.

Two accommodations are needed to execute the reference here:

* PyTables is not installed, and the reference imports `tables` at module top
  (`src/gbrs/emase/AlignmentPropertyMatrix.py:8`, `src/gbrs/emase/Sparse3DMatrix.py:13`).  `FakeTables`
  below is a ~60-line pickle-backed stand-in exposing exactly the calls the reference makes.
* For models 1-3 the reference calls `np.divide(sparse, sparse)`
  (`src/gbrs/emase/AlignmentPropertyMatrix.py:327,352,366`), which on scipy >= 1.1x returns a dense NaN
  matrix and crashes.  `NPShim` replaces that single operation with the evidently intended
  element-wise division on the numerator's sparsity pattern.  Model 4 runs
  with or without the shim, bit-identically.
"""
from __future__ import annotations

import contextlib
import io
import os
import pickle
import re
import sys
import types

import numpy as np
import scipy.sparse as sp

# the mounted reference (build container), else the copy oracle/make_ref.py staged into the git-ignored oracle/_ref/
# (that copy travels to the GPU box, where bench.py --impl reference times the unmodified reference)
_CANDIDATES = tuple(c for c in (os.environ.get("GBRS_REFERENCE_SRC"), "/root/reference/src",
                                os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "src")) if c)
REFERENCE_SRC = next((c for c in _CANDIDATES if os.path.isdir(os.path.join(c, "gbrs", "emase"))), _CANDIDATES[0])


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_SRC, "gbrs", "emase"))


# ----------------------------------------------------------------------------------------------
# fake `tables`
# ----------------------------------------------------------------------------------------------
def _norm(path: str) -> str:
    return re.sub(r"/+", "/", "/" + path.strip()) if path.strip("/") else "/"


class _Node:
    def __init__(self, fh, path):
        self._fh, self._path = fh, _norm(path)

    def read(self):
        return self._fh._store["nodes"][self._path]


class _File:
    def __init__(self, filename, mode="r", title=None):
        self._filename, self._mode = filename, mode
        if mode == "w" or not os.path.exists(filename):
            if mode == "r":
                raise FileNotFoundError(filename)
            self._store = {"attrs": {}, "nodes": {}, "title": title}
        else:
            with open(filename, "rb") as fh:
                self._store = pickle.load(fh)
        self.root = _Node(self, "/")

    @staticmethod
    def _p(where, name=None):
        base = where._path if isinstance(where, _Node) else str(where)
        return _norm(base + ("/" + name if name else ""))

    def get_node_attr(self, where, name):
        try:
            return self._store["attrs"][self._p(where)][name]
        except KeyError:
            raise AttributeError(name) from None

    def set_node_attr(self, where, name, value):
        self._store["attrs"].setdefault(self._p(where), {})[name] = value

    def get_node(self, where, name=None):
        return _Node(self, self._p(where, name))

    def __contains__(self, path):
        p = _norm(path)
        return p in self._store["nodes"] or p in self._store["attrs"]

    def create_group(self, where, name, title=None):
        p = self._p(where, name)
        self._store["attrs"].setdefault(p, {})
        return _Node(self, p)

    def create_carray(self, where, name, obj=None, title=None, filters=None):
        arr = np.asarray(obj)
        if arr.dtype.kind == "U":  # PyTables stores str arrays as bytes
            arr = arr.astype("S")
        self._store["nodes"][self._p(where, name)] = arr
        return _Node(self, self._p(where, name))

    def flush(self):
        if self._mode != "r":
            with open(self._filename, "wb") as fh:
                pickle.dump(self._store, fh)

    def close(self):
        self.flush()


def _fake_tables_module():
    m = types.ModuleType("tables")
    m.open_file = lambda filename, mode="r", title=None, **kw: _File(filename, mode, title)
    m.Filters = lambda **kw: None
    return m


# ----------------------------------------------------------------------------------------------
# numpy proxy for models 1-3
# ----------------------------------------------------------------------------------------------
class NPShim:
    def __getattr__(self, k):
        return getattr(np, k)

    @staticmethod
    def divide(a, b, *args, **kw):
        if sp.issparse(a) and sp.issparse(b):
            a = sp.csc_matrix(a)
            a.sort_indices()
            b = sp.csc_matrix(b)
            b.sort_indices()
            den = _lookup_on_pattern(a, b)
            out = sp.csc_matrix((a.data / den, a.indices.copy(), a.indptr.copy()), shape=a.shape)
            return out
        return np.divide(a, b, *args, **kw)


def _lookup_on_pattern(a: sp.csc_matrix, b: sp.csc_matrix) -> np.ndarray:
    """values of b at the stored positions of a (pattern(a) must be a subset of pattern(b))."""
    nrow = a.shape[0]
    col_a = np.repeat(np.arange(a.shape[1], dtype=np.int64), np.diff(a.indptr))
    col_b = np.repeat(np.arange(b.shape[1], dtype=np.int64), np.diff(b.indptr))
    key_a = col_a * nrow + a.indices
    key_b = col_b * nrow + b.indices
    pos = np.searchsorted(key_b, key_a)
    if pos.size and (pos.max() >= key_b.size or not np.array_equal(key_b[pos], key_a)):
        raise FloatingPointError("divide by a structural zero")
    return b.data[pos]


# ----------------------------------------------------------------------------------------------
# import + drive the reference
# ----------------------------------------------------------------------------------------------
_mods = None


def load_reference():
    """Import the reference's emase modules (with the fake `tables`).  Returns a namespace."""
    global _mods
    if _mods is not None:
        return _mods
    if not reference_available():
        raise RuntimeError("reference sources not present at " + REFERENCE_SRC)
    sys.modules.setdefault("tables", _fake_tables_module())
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)
    import gbrs.emase.AlignmentPropertyMatrix as apm_mod
    import gbrs.emase.EMfactory as em_mod
    try:  # the workflow functions are only needed for the end-to-end goldens (not staged into oracle/_ref)
        import gbrs.gbrs.emase_utils as gutils
    except ImportError:
        gutils = None

    _mods = types.SimpleNamespace(apm_mod=apm_mod, em_mod=em_mod, gutils=gutils,
                                  APM=apm_mod.AlignmentPropertyMatrix, EMfactory=em_mod.EMfactory)
    return _mods


def install_shim(enable: bool = True):
    ref = load_reference()
    ref.apm_mod.np = NPShim() if enable else np


def build_reference_apm(d, masked: bool = False):
    """Reference `AlignmentPropertyMatrix` holding SynthData `d` (SURVEY.md 8c recipe A)."""
    from gbrs_b200 import synth

    ref = load_reference()
    apm = ref.APM(shape=(d.T, d.H, d.N), haplotype_names=list(d.hname), locus_names=list(d.lname))
    mats = synth.to_csc_list(d)
    for h in range(d.H):
        apm.data[h] = mats[h]
    apm.finalize()
    apm.count = d.count.copy()
    apm.gname = np.array(d.gname)
    apm.groups = d.groups()
    apm.num_groups = len(d.gname)
    if masked:
        gm = synth.genotype_mask(d)
        apm.multiply(gm, axis=2)
        for h in range(d.H):
            apm.data[h].eliminate_zeros()
    return apm


def run_reference_em(d, model: int, lenfile: str | None, pseudocount: float = 0.0, tol: float = 1e-4,
                     max_iters: int = 999, masked: bool = False, read_length: int = 100):
    """prepare() + run() of the reference on SynthData.  Returns dict with theta0, theta, counts,
    iteration count and the per-iteration err_sum values printed by the reference."""
    ref = load_reference()
    install_shim(model != 4)
    old = np.geterr()
    try:
        apm = build_reference_apm(d, masked=masked)
        em = ref.EMfactory(apm)
        em.prepare(pseudocount=pseudocount, lenfile=lenfile, read_length=read_length)
        theta0 = np.asarray(em.allelic_expression).copy()
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            em.run(model=model, tol=tol, max_iters=max_iters, verbose=True)
        errs = [float(m.group(1)) for m in re.finditer(r"([0-9.]+) / 1000000", buf.getvalue())]
        theta = np.asarray(em.allelic_expression).copy()
        counts = np.asarray(em.probability.sum(axis=ref.APM.Axis.READ)).copy()
    finally:
        np.seterr(**old)
        install_shim(False)
    return dict(theta0=theta0, theta=theta, counts=counts, iters=len(errs), errs_printed=errs, em=em)
