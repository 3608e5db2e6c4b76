"""Small host helpers shared by the container, the EM driver and the workflow.

Logging mirrors the reference's `gbrs.utils` (/root/reference/src/gbrs/utils.py:13-90): logger name 'gbrs',
`-v` count -> WARNING / INFO / DEBUG.
"""
from __future__ import annotations

import logging
import os

import numpy as np


def get_logger(logger_name: str = "gbrs") -> logging.Logger:
    return logging.getLogger(logger_name)


def configure_logging(logger_name: str = "gbrs", level: int = 0) -> logging.Logger:
    """utils.configure_logging (utils.py:25-90): 0 -> WARNING, 1 -> INFO, 2+ -> DEBUG."""
    log = logging.getLogger(logger_name)
    try:
        from rich.logging import RichHandler

        handler = RichHandler(level=logging.NOTSET, show_level=False, show_time=True, show_path=False,
                              omit_repeated_times=False)
    except Exception:  # rich missing: plain stderr handler
        handler = logging.StreamHandler()
    debug_env = os.environ.get("GBRS_APP_DEBUG", "0") not in ("0", "", None)
    fmt = "%(message)s" if not debug_env else "[%(name)s:%(lineno)d] %(message)s"
    handler.setFormatter(logging.Formatter(fmt))
    for h in list(log.handlers):
        log.removeHandler(h)
    log.addHandler(handler)
    if level <= 0:
        log.setLevel(logging.WARNING)
    elif level == 1:
        log.setLevel(logging.INFO)
    else:
        log.setLevel(logging.DEBUG)
    log.propagate = False
    return log


def is_comment(s: str) -> bool:
    """utils.is_comment (utils.py:147-157)."""
    return s.startswith("#")


_group_cache: dict = {}  # id(groups) -> (groups, num_loci, {"gene_index": ..., "grp_conv_mat": ...})


def _cached(kind: str, num_loci: int, groups, build):
    """Gene tables are derived from the `groups` list alone; the samples of a cohort share one list object, and walking
    its ~10^5 python ints costs more than a whole EM.  One entry per live list object (the list is kept referenced, so its
    id cannot be reused)."""
    if groups is None:
        return build()
    key = id(groups)
    ent = _group_cache.get(key)
    if ent is None or ent[0] is not groups or ent[1] != num_loci or len(groups) != ent[2]:
        if len(_group_cache) >= 8:
            _group_cache.pop(next(iter(_group_cache)))
        ent = (groups, num_loci, len(groups), {})
        _group_cache[key] = ent
    if kind not in ent[3]:
        ent[3][kind] = build()
    return ent[3][kind]


def gene_index(num_loci: int, groups) -> np.ndarray:
    """int32 gene id per locus.  Loci listed in `groups[g]` get id g; loci in no group get unique ids
    >= len(groups) (they are their own singleton in the reference's `t2t_mat`, EMfactory.py:48-59)."""
    return _cached("gene_index", num_loci, groups, lambda: _gene_index(num_loci, groups)).copy()


def _gene_index(num_loci: int, groups) -> np.ndarray:
    g = np.full(num_loci, -1, dtype=np.int64)
    n_groups = len(groups) if groups is not None else 0
    if n_groups:
        sizes = np.fromiter((len(x) for x in groups), dtype=np.int64, count=n_groups)
        flat = np.fromiter((t for x in groups for t in x), dtype=np.int64, count=int(sizes.sum()))
        gid = np.repeat(np.arange(n_groups, dtype=np.int64), sizes)
        if flat.size and np.unique(flat).size != flat.size:
            # a locus repeated inside one group is harmless (the reference assigns 1.0 twice); in two groups it is not
            pairs = np.unique(np.stack((flat, gid)), axis=1)
            if np.unique(pairs[0]).size != pairs.shape[1]:
                raise NotImplementedError("a transcript listed in more than one gene group is not supported")
        g[flat] = gid
    free = np.flatnonzero(g < 0)
    g[free] = n_groups + np.arange(free.size)
    return g.astype(np.int32)


def group_conversion_matrix(num_loci: int, groups):
    """T x G 0/1 CSC matrix (`grp_conv_mat`, EMfactory.py:42-47), built without the per-gene python loop."""
    return _cached("grp_conv_mat", num_loci, groups, lambda: _group_conversion_matrix(num_loci, groups)).copy()


def _group_conversion_matrix(num_loci: int, groups):
    from scipy.sparse import csc_matrix

    n_groups = len(groups)
    sizes = np.fromiter((len(x) for x in groups), dtype=np.int64, count=n_groups)
    rows = np.fromiter((t for x in groups for t in x), dtype=np.int64, count=int(sizes.sum()))
    indptr = np.zeros(n_groups + 1, dtype=np.int64)
    indptr[1:] = np.cumsum(sizes)
    m = csc_matrix((np.ones(rows.size), rows, indptr), shape=(num_loci, n_groups))
    m.sum_duplicates()
    m.data[:] = 1.0
    return m


def write_table_rows(fh, names, cntdata: np.ndarray, notes=None, order=None) -> None:
    """Rows `name<TAB>v0<TAB>v1...[<TAB>note]` with values formatted like the reference's `str(numpy.float64)`
    (EMfactory.py:325-331, :370-380).  The formatting runs natively (libgbrs_em.so, all host threads): at 80k loci a
    table is ~0.7 M numbers and the python loop took longer than the whole EM.  `fh` is an open text file positioned
    after the header; the rows are appended to the same file."""
    import ctypes as C

    from . import _lib

    lib = _lib.load()
    n = len(names)
    data = np.ascontiguousarray(cntdata, dtype=np.float64)
    if data.ndim != 2 or data.shape[1] != n:
        raise ValueError("cntdata must be [columns][rows]")
    c_names = (C.c_char_p * n)(*[str(x).encode() for x in names])
    c_notes = None
    if notes is not None:
        c_notes = (C.c_char_p * n)(*[str(notes[x]).encode() for x in names])
    c_order, n_rows = None, n
    if order is not None:
        order = np.ascontiguousarray(order, dtype=np.int64)
        c_order, n_rows = order.ctypes.data, int(order.shape[0])
    fh.flush()
    path = fh.name
    # rows index `names` / `data` through `order`; n_rows of them are written, the arrays keep their full length
    if c_order is not None and n_rows != n:
        raise ValueError("order must be a permutation of the rows")
    _lib.check(lib.gbrs_write_table(path.encode(), None, c_names, n, data.ctypes.data, data.shape[0], c_notes, c_order, 1))
    fh.seek(0, 2)


def py_float_str(x: float) -> str:
    """The native formatter for one value (tests compare it with python's repr)."""
    import ctypes as C

    from . import _lib

    buf = C.create_string_buffer(64)
    n = _lib.load().gbrs_format_double(float(x), buf, 64)
    if n < 0:
        raise ValueError("formatting failed")
    return buf.value.decode()


def read_npz_members(path: str, names=None, workers: int | None = None) -> dict:
    """name -> array for members of an `.npz` file, without numpy's per-member overhead.  `names`: a list of member
    names (KeyError for one the file lacks), a predicate on the member name, or None for all members.

    `np.load(path)[name]` opens the member, parses its header with `ast.literal_eval` and inflates it, one member at a
    time: ~0.15 ms for each of the 22k small matrices of an alignment-specificity file and 2.4 s of single-threaded zlib
    for the 20 large matrices of a transition file.  Here large members are dealt over a pool of threads (zlib releases
    the GIL), every thread with a zip handle of its own; small members are interpreter-bound and stay on one thread.  A
    header is parsed only when its bytes differ from the previous member's (same dtype and shape: the common case).
    Fortran-ordered and object arrays fall back to numpy."""
    import io
    import zipfile
    from concurrent.futures import ThreadPoolExecutor

    def load_chunk(chunk, zf):
        out = {}
        last_hdr, last_meta = None, None
        for name in chunk:
            raw = zf.read(name + ".npy")
            if raw[:6] != b"\x93NUMPY":
                raise ValueError(f"{path}: member {name} is not an .npy array")
            major = raw[6]
            if major == 1:
                hlen, off = int.from_bytes(raw[8:10], "little"), 10
            elif major == 2:
                hlen, off = int.from_bytes(raw[8:12], "little"), 12
            else:  # format 3 (utf-8 field names) and anything newer: numpy's own reader
                out[name] = np.load(io.BytesIO(raw), allow_pickle=False)
                continue
            hdr = raw[off: off + hlen]
            if hdr != last_hdr:
                fh = io.BytesIO(raw)
                np.lib.format.read_magic(fh)
                shape, fortran, dtype = (np.lib.format.read_array_header_1_0(fh) if major == 1
                                         else np.lib.format.read_array_header_2_0(fh))
                last_hdr, last_meta = hdr, (shape, fortran, dtype)
            shape, fortran, dtype = last_meta
            if fortran or dtype.hasobject:
                out[name] = np.load(io.BytesIO(raw), allow_pickle=False)
            else:
                out[name] = np.frombuffer(raw, dtype=dtype, offset=off + hlen,
                                          count=int(np.prod(shape, dtype=np.int64))).reshape(shape)
        return out

    def load_chunk_own_handle(chunk):
        with zipfile.ZipFile(path) as zf:
            return load_chunk(chunk, zf)

    with zipfile.ZipFile(path) as zf0:
        if names is None or callable(names):
            have = [n[:-4] for n in zf0.namelist() if n.endswith(".npy")]
            want = have if names is None else [n for n in have if names(n)]
        else:
            want = list(names)
        if not want:
            return {}
        if workers is None:
            # small members are interpreter-bound (zip bookkeeping, not zlib): threads would only fight over the GIL
            # and re-read the zip directory once each
            small = os.path.getsize(path) / max(len(zf0.namelist()), 1) < (1 << 18)
            workers = 1 if small else (os.cpu_count() or 1)
        workers = max(1, min(workers, 16, len(want)))
        if workers == 1:
            return load_chunk(want, zf0)
    merged = {}
    with ThreadPoolExecutor(max_workers=workers) as pool:
        for part in pool.map(load_chunk_own_handle, [want[w::workers] for w in range(workers)]):
            merged.update(part)
    return {n: merged[n] for n in want}
