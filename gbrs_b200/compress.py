"""`compress` -- equivalence-class construction and replicate merging, the step right before `quantify`.

Same call surface and result as the reference's `compress` (/root/reference/src/gbrs/gbrs/emase_utils.py:22-107):
every read of every input EMASE file (in input order) is keyed by the set of (locus, haplotype) positions it hits;
reads with equal keys are merged into one alignment class whose count is the sum of the reads' counts (all ones when a
file carries no count vector, :58-59); classes are numbered in order of first appearance (the reference's dict
insertion order, :72, :92) and written as a compressed-EMASE incidence matrix with class counts.

The reference builds a Python string per read and a dict; here the grouping runs on the GPU through the C ABI
(`gbrs_ec_build`, gbrs_b200/csrc/ec_kernels.cu: hash, radix sort, exact comparison, first-appearance numbering).  The
host side only re-lays the H per-haplotype sparse matrices as one row of (locus | mask << 24) words per read
(`gbrs_rows_create`, native and threaded) and expands the representative rows back into H CSC matrices.  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
from scipy.sparse import csc_matrix

from . import _lib, utils
from .apm import AlignmentPropertyMatrix

logger = utils.get_logger("gbrs")

_SEEDS = (0x243F6A8885A308D3, 0x13198A2E03707344, 0xA4093822299F31D0, 0x082EFA98EC4E6C89)


def read_rows(mats, T: int, H: int):
    """H sparse matrices (reads x loci) -> CSR of pair words per read: rowptr [n+1] (int64), words [pairs] (uint32,
    locus | hapmask << 24, ascending locus inside a read).  Stored zeros are not alignments and are dropped.  Native
    (gbrs_rows_create: threaded merge + transposition); the result is copied out and the C object freed."""
    if H > _lib.GBRS_HPAD:
        raise NotImplementedError("more than 8 haplotypes is not supported by the packed mask layout")
    if T >= (1 << 24):
        raise NotImplementedError("more than 2^24 loci is not supported by the packed pair words")
    lib = _lib.load()
    n = mats[0].shape[0]
    csc = [m if m.format == "csc" else m.tocsc() for m in mats]
    indptr = [np.ascontiguousarray(m.indptr, dtype=np.int64) for m in csc]
    wide = any(m.indices.dtype.itemsize == 8 for m in csc)
    indices = [np.ascontiguousarray(m.indices, dtype=np.int64 if wide else np.int32) for m in csc]
    values = [np.ascontiguousarray(m.data, dtype=np.float64) for m in csc]
    inp = _lib.PackInput()
    inp.T, inp.H, inp.N = T, H, n
    inp.indptr = (C.c_void_p * H)(*[a.ctypes.data for a in indptr])
    inp.indices = (C.c_void_p * H)(*[a.ctypes.data for a in indices])
    inp.values = (C.c_void_p * H)(*[a.ctypes.data for a in values])
    inp.index_bytes = 8 if wide else 4
    inp.shard_rank, inp.shard_count = 0, 1
    handle = C.c_void_p()
    _lib.check(lib.gbrs_rows_create(C.byref(inp), C.byref(handle)))
    try:
        p_rowptr, p_words, n_words = C.c_void_p(), C.c_void_p(), C.c_int64()
        _lib.check(lib.gbrs_rows_get(handle, C.byref(p_rowptr), C.byref(p_words), C.byref(n_words)))
        rowptr = np.frombuffer((C.c_int64 * (n + 1)).from_address(p_rowptr.value), dtype=np.int64).copy()
        if n_words.value:
            words = np.frombuffer((C.c_uint32 * n_words.value).from_address(p_words.value), dtype=np.uint32).copy()
        else:
            words = np.zeros(0, dtype=np.uint32)
    finally:
        lib.gbrs_rows_free(handle)
    return rowptr, words


def equivalence_classes(rowptr, words, count=None, device=None):
    """Group identical rows on the GPU.  Returns (class_of_read [n] uint32, first_read [n_classes] uint32,
    class_count [n_classes] float64); classes are numbered by first appearance."""
    import torch

    lib = _lib.load()
    if not torch.cuda.is_available():
        raise _lib.GbrsCudaError("no CUDA device is available; gbrs_b200 compress has no CPU fallback")
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    n = int(rowptr.shape[0]) - 1
    if n >= (1 << 32) or int(rowptr[-1]) >= (1 << 32):
        raise NotImplementedError("more than 2^32 reads / pair words in one compress call")
    if n == 0:
        return np.zeros(0, np.uint32), np.zeros(0, np.uint32), np.zeros(0)
    with torch.cuda.device(dev):
        d_rowptr = torch.from_numpy(rowptr.astype(np.uint32).view(np.int32)).to(dev)
        d_words = torch.from_numpy(np.ascontiguousarray(words, dtype=np.uint32).view(np.int32)).to(dev)
        if d_words.numel() == 0:
            d_words = torch.zeros(4, dtype=torch.int32, device=dev)
        d_count = None if count is None else torch.from_numpy(np.ascontiguousarray(count, dtype=np.float64)).to(dev)
        cls = torch.empty(n, dtype=torch.int32, device=dev)
        first = torch.empty(n, dtype=torch.int32, device=dev)
        ccount = torch.empty(n, dtype=torch.float64, device=dev)
        nbytes = C.c_int64()
        _lib.check(lib.gbrs_ec_workspace_bytes(n, C.byref(nbytes)))
        work = torch.empty(nbytes.value, dtype=torch.uint8, device=dev)
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        n_classes, collisions = C.c_int64(), C.c_int64()
        for seed in _SEEDS:
            _lib.check(lib.gbrs_ec_build(n, d_rowptr.data_ptr(), d_words.data_ptr(),
                                         None if d_count is None else d_count.data_ptr(), seed, cls.data_ptr(),
                                         first.data_ptr(), ccount.data_ptr(), work.data_ptr(), nbytes.value, stream,
                                         C.byref(n_classes), C.byref(collisions)))
            if collisions.value == 0:
                break
            logger.debug(f"hash collision between different alignment patterns ({collisions.value}); new seed")
        else:
            raise _lib.GbrsError("could not separate the alignment patterns by hashing (four seeds collided)")
        k = int(n_classes.value)
        return (cls.cpu().numpy().view(np.uint32), first[:k].cpu().numpy().view(np.uint32), ccount[:k].cpu().numpy())


def class_matrices(rowptr, words, first_read, T: int, H: int):
    """Rows of the representative reads -> H CSC matrices (classes x loci), values 1.0."""
    k = int(first_read.shape[0])
    lens = (rowptr[1:] - rowptr[:-1])[first_read]
    starts = rowptr[:-1][first_read]
    total = int(lens.sum())
    cls = np.repeat(np.arange(k, dtype=np.int64), lens)
    off = np.arange(total, dtype=np.int64) - np.repeat(np.cumsum(lens) - lens, lens)
    w = words[np.repeat(starts, lens) + off]
    locus, mask = (w & np.uint32(0xFFFFFF)).astype(np.int64), (w >> np.uint32(24)).astype(np.int64)
    mats = []
    for h in range(H):
        sel = ((mask >> h) & 1).astype(bool)
        m = csc_matrix((np.ones(int(sel.sum())), (cls[sel], locus[sel])), shape=(k, T))
        m.sort_indices()
        mats.append(m)
    return mats


def compress(emase_files: list[str], output_file: str, comp_lib: str = "zlib", device=None) -> None:
    """Compress EMASE file(s) to an alignment incidence matrix of equivalence classes (emase_utils.py:22-107)."""
    for x in emase_files:
        logger.info(f"EMASE file: {x}")
    logger.info(f"Output File: {output_file}")
    logger.info(f"Compression Library: {comp_lib}")
    num_loci = num_haplotypes = names_loci = names_haplotypes = None
    rowptrs, words, counts = [], [], []
    base = 0
    for aln_file in emase_files:
        logger.info(f"Loading EMASE file: {aln_file}")
        apm = AlignmentPropertyMatrix(h5file=aln_file)
        logger.debug(f"Number Loci: {apm.num_loci}")
        logger.debug(f"Number Haplotypes: {apm.num_haplotypes}")
        logger.debug(f"Number Reads: {apm.num_reads}")
        if num_loci is not None and (apm.num_loci, apm.num_haplotypes) != (num_loci, num_haplotypes):
            raise RuntimeError("the EMASE files to merge differ in their number of loci / haplotypes")
        # each file should be the same (:54-57)
        num_loci, num_haplotypes = apm.num_loci, apm.num_haplotypes
        names_loci, names_haplotypes = apm.lname, apm.hname
        rp, w = read_rows(apm.data, num_loci, num_haplotypes)
        rowptrs.append(rp[1:] + base if rowptrs else rp)
        base += int(rp[-1])
        words.append(w)
        counts.append(np.ones(apm.num_reads) if apm.count is None else np.asarray(apm.count, dtype=np.float64))  # :58-59
    if num_loci is None:
        raise RuntimeError("no EMASE file given")
    rowptr = np.concatenate(rowptrs)
    words = np.concatenate(words)
    count = np.concatenate(counts)
    logger.debug("Creating unique ECs")
    _, first_read, ec_count = equivalence_classes(rowptr, words, count, device=device)
    num_ecs = int(first_read.shape[0])
    logger.info("Constructing APM")
    logger.debug(f"Number Loci: {num_loci}")
    logger.debug(f"Number Haplotypes: {num_haplotypes}")
    logger.debug(f"Number ECs: {num_ecs}")
    mats = class_matrices(rowptr, words, first_read, num_loci, num_haplotypes)
    out = AlignmentPropertyMatrix.from_csc(mats, names_haplotypes, names_loci, count=ec_count)
    logger.info(f"Saving EMASE Formatted File: {output_file}")
    out.save(h5file=output_file, complib=comp_lib)
    logger.info("Done")
