"""Host-side container for a compressed-EMASE alignment incidence matrix.

Mirrors the fields and constructor of the reference's `AlignmentPropertyMatrix`
(/root/reference/src/gbrs/emase/AlignmentPropertyMatrix.py:25-111) and its base `Sparse3DMatrix`
(src/gbrs/emase/Sparse3DMatrix.py:26-66): `shape = (T loci, H haplotypes, N classes)`, `data` = list of H
scipy CSC matrices (N x T), `count`, `hname`, `lname`, `lid`, `gname`, `groups`, `num_groups`.

In this implementation the object is only a *container*: the numerical methods of the reference that form
the EM inner loop (`reset`, `multiply` along READ/HAPLOTYPE, `normalize_reads`, `sum(READ)`) are not executed
on the host at all -- `gbrs_b200.EMfactory` packs the pattern once and runs them fused on the GPU.  What is
kept here is what `quantify` needs around the loop: loading/saving, group loading, the `-G` masking entry
point (`multiply(gtmask, axis=2)`), and `report_alignment_counts`.

File formats.  The reference stores the matrix with PyTables (HDF5).  PyTables/HDF5 are not available in the
build image, so two formats are accepted: the reference's HDF5 layout when `tables` is importable, and an
`.npz` twin with the same logical fields (`shape`, `hname`, `lname`, `count`, `h{k}_indptr`, `h{k}_indices`,
optional `h{k}_data`, `incidence_only`).  `save()` writes whichever the environment supports (HDF5 only when
PyTables is present and the name does not end in `.npz`).
"""
from __future__ import annotations

import os

import copy
from enum import IntEnum

import numpy as np
from scipy.sparse import csc_matrix, lil_matrix

from . import utils

logger = utils.get_logger("gbrs")

_HDF5_MAGIC = b"\x89HDF\r\n\x1a\n"


class AxisEnum(IntEnum):  # AlignmentPropertyMatrix.py:17-22
    LOCUS = 0
    HAPLOTYPE = 1
    READ = 2
    GROUP = 3
    HAPLOGROUP = 4


def _try_import_tables():
    try:
        import tables  # noqa: F401

        return tables
    except Exception:
        return None


class AlignmentPropertyMatrix:
    Axis = AxisEnum

    def __init__(self, other=None, h5file=None, datanode="/", metanode="/", shallow=False, shape=None,
                 dtype=float, haplotype_names=None, locus_names=None, read_names=None, grpfile=None):
        self.shape = (0, 0, 0)
        self.ndim = 3
        self.data = []
        self.finalized = False
        self.num_groups = 0
        self.count = None
        self.hname = None
        self.lname = None
        self.rname = None
        self.lid = None
        self.rid = None
        self.gname = None
        self.groups = None

        if other is not None:
            if not other.finalized:
                raise RuntimeError("The original matrix must be finalized.")
            self.shape = other.shape
            self.data = copy.deepcopy(other.data)
            self.finalized = True
            if other.count is not None:
                self.count = copy.copy(other.count)
            if not shallow:
                self._copy_names(other)
                self._copy_group_info(other)
        elif h5file is not None:
            self._load(str(h5file), datanode, metanode, shallow, dtype)
        elif shape is not None:
            if len(shape) != 3 or (np.array(shape) < 1).any():
                raise RuntimeError("The shape must be a tuple of three positive integers.")
            self.shape = tuple(int(x) for x in shape)
            for _ in range(self.shape[1]):
                self.data.append(lil_matrix((self.shape[2], self.shape[0]), dtype=dtype))
        self.num_loci, self.num_haplotypes, self.num_reads = self.shape
        if other is None and h5file is None and shape is not None:
            if haplotype_names is not None:
                if len(haplotype_names) != self.num_haplotypes:
                    raise RuntimeError("The number of names does not match to the matrix shape.")
                self.hname = haplotype_names
            if locus_names is not None:
                if len(locus_names) != self.num_loci:
                    raise RuntimeError("The number of names does not match to the matrix shape.")
                self.lname = np.array(locus_names)
                self.lid = dict(zip(self.lname, np.arange(self.num_loci)))
            if read_names is not None:
                if len(read_names) != self.num_reads:
                    raise RuntimeError("The number of names does not match to the matrix shape.")
                self.rname = np.array(read_names)
                self.rid = dict(zip(self.rname, np.arange(self.num_reads)))
        if grpfile is not None:
            self.load_groups(grpfile)

    @classmethod
    def from_csc(cls, mats, haplotype_names, locus_names, count=None):
        """Wrap H ready-made scipy CSC matrices (N x T) without the lil_matrix detour of the `shape=` constructor."""
        self = cls()
        N, T = mats[0].shape
        self.shape = (int(T), len(mats), int(N))
        self.num_loci, self.num_haplotypes, self.num_reads = self.shape
        self.data = [m.tocsc() for m in mats]
        self.finalized = True
        self.hname = list(haplotype_names)
        if len(self.hname) != self.num_haplotypes or len(locus_names) != self.num_loci:
            raise RuntimeError("The number of names does not match to the matrix shape.")
        self.lname = np.array(locus_names)
        self.lid = dict(zip(self.lname, np.arange(self.num_loci)))
        if count is not None:
            self.count = np.asarray(count, dtype=np.float64)
        return self

    # ------------------------------------------------------------------ loading / saving
    def _load(self, path, datanode, metanode, shallow, dtype):
        with open(path, "rb") as fh:
            magic = fh.read(8)
        if magic == _HDF5_MAGIC:
            self._load_hdf5(path, datanode, metanode, shallow, dtype)
        else:
            self._load_npz(path, shallow, dtype)
        self.finalized = True

    def _load_npz(self, path, shallow, dtype):
        """The `.npz` twin of the PyTables layout.  The per-haplotype arrays are inflated and turned into CSC matrices
        by a pool of threads (zlib and numpy's fills release the GIL): at benchmark scale the single-threaded inflate
        was the largest part of a whole `quantify` run."""
        from concurrent.futures import ThreadPoolExecutor

        with np.load(path, allow_pickle=False) as z:
            T, H, N = (int(x) for x in z["shape"])
            self.shape = (T, H, N)
            files = set(z.files)
            incidence_only = bool(z["incidence_only"]) if "incidence_only" in files else True

            def one(h):
                with np.load(path, allow_pickle=False) as zz:  # a handle of its own: no shared file position
                    indptr = zz[f"h{h}_indptr"]
                    indices = zz[f"h{h}_indices"]
                    if not incidence_only and f"h{h}_data" in files:
                        vals = zz[f"h{h}_data"].astype(dtype)
                    else:
                        vals = np.ones(indices.shape[0], dtype=dtype)
                idx_t = np.int64 if max(N, indices.shape[0]) >= 2**31 - 1 else np.int32
                return csc_matrix((vals, indices.astype(idx_t), indptr.astype(idx_t)), shape=(N, T))

            workers = max(1, min(H, os.cpu_count() or 1))
            if workers > 1:
                with ThreadPoolExecutor(max_workers=workers) as pool:
                    self.data = list(pool.map(one, range(H)))
            else:
                self.data = [one(h) for h in range(H)]
            if "count" in files:
                self.count = z["count"].astype(np.float64)
            if not shallow:
                self.hname = [str(x) for x in z["hname"]]
                self.lname = [str(x) for x in z["lname"]]
                self.lid = dict(zip(self.lname, np.arange(T)))
                if "rname" in files:
                    self.rname = z["rname"]
                    self.rid = dict(zip(self.rname, np.arange(N)))

    def _load_hdf5(self, path, datanode, metanode, shallow, dtype):
        """The reference's PyTables layout (Sparse3DMatrix.py:42-50, :68-102; AlignmentPropertyMatrix.py:70-83)."""
        tables = _try_import_tables()
        if tables is None:
            raise RuntimeError(
                f"{path} is an HDF5 (PyTables) EMASE file but PyTables is not installed in this environment; "
                "convert it where PyTables exists with `python -m gbrs_b200.convert in.h5 out.npz`")
        h5fh = tables.open_file(path, "r")
        try:
            self.shape = tuple(int(x) for x in h5fh.get_node_attr(datanode, "shape"))
            T, H, N = self.shape
            try:
                mtype = h5fh.get_node_attr(datanode, "mtype")
                mtype = mtype.decode() if isinstance(mtype, bytes) else mtype
                incidence_only = h5fh.get_node_attr(datanode, "incidence_only")
            except AttributeError:
                mtype, incidence_only = "coo_matrix", False
            for h in range(H):
                node = h5fh.get_node(f"{datanode}/h{h}")
                if mtype == "csc_matrix":
                    indptr = h5fh.get_node(node, "indptr").read().astype(np.int64)
                    indices = h5fh.get_node(node, "indices").read().astype(np.int64)
                    vals = (np.ones(len(indices), dtype=dtype) if incidence_only
                            else h5fh.get_node(node, "data").read().astype(dtype))
                    self.data.append(csc_matrix((vals, indices, indptr), shape=(N, T)))
                elif mtype == "coo_matrix":
                    from scipy.sparse import coo_matrix

                    coor = h5fh.get_node(node, "coor").read()
                    vals = h5fh.get_node(node, "data").read().astype(dtype)
                    self.data.append(coo_matrix((vals, coor), shape=(N, T)).tocsc())
                else:
                    raise RuntimeError("Only csc or coo matrices are supported.")
            if f"{datanode}/count" in h5fh:
                self.count = h5fh.get_node(datanode, "count").read()
            if not shallow:
                self.hname = h5fh.get_node_attr(datanode, "hname")
                self.lname = [x.decode() for x in h5fh.get_node(metanode, "lname").read()]
                self.lid = dict(zip(self.lname, np.arange(T)))
                if f"{metanode}/rname" in h5fh:
                    self.rname = h5fh.get_node(metanode, "rname").read()
                    self.rid = dict(zip(self.rname, np.arange(N)))
        finally:
            h5fh.close()

    def save(self, h5file, title=None, index_dtype="uint32", data_dtype=float, incidence_only=True,
             complib="zlib", shallow=False):
        """AlignmentPropertyMatrix.save (AlignmentPropertyMatrix.py:478-525): pattern (+ values unless
        `incidence_only`), counts and names."""
        if not self.finalized:
            raise RuntimeError("The matrix is not finalized.")
        tables = _try_import_tables()
        if tables is not None and not str(h5file).endswith(".npz"):
            self._save_hdf5(tables, h5file, title, index_dtype, data_dtype, incidence_only, complib, shallow)
            return
        out = {"shape": np.array(self.shape, dtype=np.int64), "incidence_only": np.array(bool(incidence_only)),
               "mtype": np.array("csc_matrix")}
        for h in range(self.shape[1]):
            m = self.data[h]
            out[f"h{h}_indptr"] = m.indptr.astype(index_dtype)
            out[f"h{h}_indices"] = m.indices.astype(index_dtype)
            if not incidence_only:
                out[f"h{h}_data"] = m.data.astype(data_dtype)
        if self.count is not None:
            out["count"] = np.asarray(self.count)
        if not shallow:
            out["hname"] = np.array(list(self.hname))
            out["lname"] = np.array(list(self.lname))
            if self.rname is not None:
                out["rname"] = np.asarray(self.rname)
        with open(h5file, "wb") as fh:  # keep the exact file name (np.savez would append .npz)
            np.savez_compressed(fh, **out)

    def _save_hdf5(self, tables, h5file, title, index_dtype, data_dtype, incidence_only, complib, shallow):
        h5fh = tables.open_file(h5file, "w", title=title)
        fil = tables.Filters(complevel=1, complib=complib)
        h5fh.set_node_attr(h5fh.root, "incidence_only", incidence_only)
        h5fh.set_node_attr(h5fh.root, "mtype", "csc_matrix")
        h5fh.set_node_attr(h5fh.root, "shape", self.shape)
        for h in range(self.shape[1]):
            grp = h5fh.create_group(h5fh.root, f"h{h}", f"Sparse matrix components for Haplotype {h}")
            m = self.data[h]
            h5fh.create_carray(grp, "indptr", obj=m.indptr.astype(index_dtype), filters=fil)
            h5fh.create_carray(grp, "indices", obj=m.indices.astype(index_dtype), filters=fil)
            if not incidence_only:
                h5fh.create_carray(grp, "data", obj=m.data.astype(data_dtype), filters=fil)
        if self.count is not None:
            h5fh.create_carray(h5fh.root, "count", obj=self.count, title="Equivalence Class Counts", filters=fil)
        if not shallow:
            h5fh.set_node_attr(h5fh.root, "hname", self.hname)
            h5fh.create_carray(h5fh.root, "lname", obj=np.array(self.lname), title="Locus Names", filters=fil)
            if self.rname is not None:
                h5fh.create_carray(h5fh.root, "rname", obj=self.rname, title="Read Names", filters=fil)
        h5fh.flush()
        h5fh.close()

    # ------------------------------------------------------------------ groups / names
    def load_groups(self, grpfile):
        """gene<TAB>t1<TAB>t2...  (AlignmentPropertyMatrix.py:113-130)."""
        if self.lid is None:
            raise RuntimeError("Locus IDs are not available.")
        gname, groups = [], []
        with open(grpfile) as fh:
            for curline in fh:
                item = curline.rstrip().split("\t")
                gname.append(item[0])
                groups.append([int(self.lid[t]) for t in item[1:]])
        self.gname = np.array(gname)
        self.groups = groups
        self.num_groups = len(gname)

    def _copy_names(self, other):
        self.hname = other.hname
        self.lname = copy.copy(other.lname)
        self.rname = copy.copy(other.rname)
        self.lid = copy.copy(other.lid)
        self.rid = copy.copy(other.rid)

    def _copy_group_info(self, other):
        if other.groups is not None and other.gname is not None:
            self.groups = copy.deepcopy(other.groups)
            self.gname = copy.copy(other.gname)
            self.num_groups = other.num_groups

    def copy(self, shallow=False):
        return AlignmentPropertyMatrix(other=self, shallow=shallow)

    def finalize(self):
        if not self.finalized:
            for h in range(self.shape[1]):
                self.data[h] = self.data[h].tocsc()
            self.finalized = True

    @property
    def nnz(self) -> int:
        return int(sum(m.nnz for m in self.data))

    # ------------------------------------------------------------------ the one in-place op quantify uses
    def multiply(self, multiplier, axis=None):
        """Only the branch `quantify -G` uses: 2-D H x T multiplier along axis=2
        (Sparse3DMatrix.py:354-362): scale every column of every haplotype matrix."""
        if not self.finalized:
            raise RuntimeError("The original matrix must be finalized.")
        self._pure_cache = None
        multiplier = np.asarray(multiplier)
        if multiplier.ndim == 2 and axis == 2:
            for h in range(self.shape[1]):
                m = self.data[h]
                m.data *= np.repeat(multiplier[h, :].ravel(), np.diff(m.indptr))
            return
        raise NotImplementedError(
            "gbrs_b200 executes the EM's multiply/normalize steps fused on the GPU (EMfactory); only the "
            "2-D axis=2 genotype-mask form is available on the host container")

    def eliminate_zeros(self):
        for h in range(self.shape[1]):
            self.data[h].eliminate_zeros()

    def is_pure_incidence(self, cache: bool = False) -> bool:
        """True if every stored value is 1.0.  `cache=True` remembers the answer on the object (the cohort loader thread
        asks ahead of time; a pass over the values of a million classes costs more than their EM) -- callers that change
        the values afterwards must not rely on it (`multiply` and `reset` drop it)."""
        if cache and getattr(self, "_pure_cache", None) is not None:
            return self._pure_cache
        pure = all(bool(np.all(m.data == 1.0)) for m in self.data)
        if cache:
            self._pure_cache = pure
        return pure

    def reset(self):
        """Sparse3DMatrix.reset (Sparse3DMatrix.py:220-228): values := 1 on the current pattern."""
        self._pure_cache = None
        if not self.finalized:
            raise RuntimeError("The original matrix must be finalized.")
        for h in range(self.shape[1]):
            self.data[h].data = np.ones(self.data[h].nnz, dtype=self.data[h].dtype)

    # ------------------------------------------------------------------ alignment counts (device)
    def _alignment_count_tables(self, gene_level=False, device=None, pattern=None):
        """(aln, uniq, locus_uniq) on the device.  `pattern`: an already packed + resident `DevicePattern` of THIS
        (unmasked) matrix to reuse -- `quantify` passes the EM's own pattern for multiway runs, so that the matrix is
        neither reloaded nor packed again; otherwise one pattern is packed on first use and kept."""
        from .emfactory import DevicePattern

        n_real = 0
        if gene_level:
            if not (self.num_groups > 0 and self.groups is not None and self.gname is not None):
                raise RuntimeError("No group information is available for bundling.")
            n_real = self.num_groups
        if pattern is None:
            pattern = getattr(self, "_count_pattern", None)
            need_genes = self.num_groups > 0 and self.groups is not None
            if pattern is None or (gene_level and not pattern.packed.has_genes):
                gene_of = utils.gene_index(self.num_loci, self.groups) if need_genes else None
                pattern = DevicePattern(self, gene_of=gene_of, device=device)
                self._count_pattern = pattern
        elif gene_level and not pattern.packed.has_genes:
            raise RuntimeError("the supplied pattern was packed without gene information")
        return pattern.alignment_counts(gene_level=gene_level, n_real_genes=n_real)

    def count_alignments(self):
        """H x T count-weighted alignment counts (AlignmentPropertyMatrix.py:436-440)."""
        return self._alignment_count_tables()[0]

    def count_unique_reads(self, ignore_haplotype=False):
        aln, uniq, locus_uniq = self._alignment_count_tables()
        return locus_uniq if ignore_haplotype else uniq

    def report_alignment_counts(self, filename, gene_level=False, pattern=None):
        """AlignmentPropertyMatrix.report_alignment_counts (:442-459).  `gene_level=True` is what the reference
        obtains by `_bundle_inline(reset=True)` followed by this call (gbrs/emase_utils.py:327-331)."""
        aln, uniq, locus_uniq = self._alignment_count_tables(gene_level=gene_level, pattern=pattern)
        names = self.gname if gene_level else self.lname
        cntdata = np.vstack((aln, uniq, locus_uniq[None, :]))
        with open(filename, "w") as fh:
            fh.write("locus\t" + "\t".join(f"aln_{h}" for h in self.hname) + "\t")
            fh.write("\t".join(f"uniq_{h}" for h in self.hname) + "\t")
            fh.write("locus_uniq\n")
            utils.write_table_rows(fh, names, cntdata)

    def _bundle_inline(self, reset=False):
        """_bundle_inline (AlignmentPropertyMatrix.py:155-188): loci -> genes on the host container."""
        if not self.finalized:
            raise RuntimeError("The matrix is not finalized.")
        if not (self.num_groups > 0 and self.groups is not None and self.gname is not None):
            raise RuntimeError("No group information is available for bundling.")
        conv = utils.group_conversion_matrix(self.num_loci, self.groups)
        for h in range(self.num_haplotypes):
            self.data[h] = (self.data[h] * conv).tocsc()
        self.num_loci = self.num_groups
        self.shape = (self.num_groups, self.num_haplotypes, self.num_reads)
        self.lname = copy.copy(self.gname)
        self.lid = dict(zip(self.gname, np.arange(self.num_groups)))
        self.num_groups = 0
        self.groups = None
        self.gname = None
        if reset:
            self.reset()
