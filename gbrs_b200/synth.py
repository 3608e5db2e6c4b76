"""Seeded synthetic compressed-EMASE alignment incidence data.

This is the canonical generator of SURVEY.md section 8(d) / BASELINE.md section 3.  It produces the
logical content of a compressed EMASE file (transcripts x haplotypes x alignment classes, with class
counts), the gene -> transcript grouping, the per-(locus, haplotype) lengths and, optionally, a diploid
genotype call per gene.  Everything is returned as plain numpy arrays in *class-major pair form*
(class id, locus id, 8-bit haplotype mask); `to_csc_list()` turns that into the reference's storage
(`list` of H scipy CSC matrices, each N_classes x T_loci, cf. reference
`src/gbrs/emase/Sparse3DMatrix.py:26-66`).

Nothing here is on the product path; tests, fixtures and bench.py use it to make inputs.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

BASE_SEED = 20261018
HAPLOTYPES = ("A", "B", "C", "D", "E", "F", "G", "H")


@dataclass
class SynthData:
    T: int
    H: int
    N: int
    pair_class: np.ndarray  # int64 [pairs], non-decreasing
    pair_locus: np.ndarray  # int64 [pairs]
    pair_mask: np.ndarray  # uint8 [pairs], never 0
    count: np.ndarray  # float64 [N], integer valued
    gene_of: np.ndarray  # int64 [T], contiguous loci per gene, dense gene ids 0..G-1
    lengths: np.ndarray  # float64 [T, H] raw transcript lengths (not yet effective lengths)
    hname: tuple = HAPLOTYPES
    lname: list = field(default_factory=list)
    gname: list = field(default_factory=list)
    genotype: list | None = None  # per gene two-letter diplotype (sorted letters) or None

    @property
    def G(self) -> int:
        return len(self.gname)

    @property
    def pairs(self) -> int:
        return int(self.pair_class.shape[0])

    @property
    def nnz(self) -> int:
        return int(_popcount8(self.pair_mask).sum())

    def groups(self) -> list[list[int]]:
        """gene -> list of locus ids (reference `AlignmentPropertyMatrix.groups`)."""
        bounds = np.flatnonzero(np.diff(self.gene_of)) + 1
        return [list(map(int, x)) for x in np.split(np.arange(self.T), bounds)]


def _popcount8(x: np.ndarray) -> np.ndarray:
    table = np.array([bin(i).count("1") for i in range(256)], dtype=np.int64)
    return table[x.astype(np.uint8)]


def generate(T: int, N: int, H: int = 8, sample_index: int = 0, with_genotype: bool = False,
             n_genes: int | None = None, wide_frac: float = 0.0, wide_max: int = 24) -> SynthData:
    """Generate one sample.  `sample_index` shifts the seed (cohort mode shares gene_of / lengths by
    drawing them from the base seed first).  `wide_frac` > 0 additionally makes that fraction of the classes
    "wide" (9..wide_max loci, spread over a window of 4*wide_max loci) to exercise the long-class code paths; the
    canonical benchmark shapes use wide_frac = 0."""
    if not (1 <= H <= 8):
        raise ValueError("synthetic generator supports 1..8 haplotypes")
    shared = np.random.default_rng(BASE_SEED)
    G0 = max(1, int(0.4 * T)) if n_genes is None else int(n_genes)
    gene_raw = np.sort(shared.integers(0, G0, T))
    # dense gene ids (drop genes that drew no locus)
    _, gene_of = np.unique(gene_raw, return_inverse=True)
    gene_of = gene_of.astype(np.int64)
    lengths = shared.integers(300, 6000, size=(T, H)).astype(np.float64)
    geno_idx = shared.integers(0, H, size=(int(gene_of.max()) + 1, 2))

    rng = np.random.default_rng(BASE_SEED + 1 + sample_index)
    k = np.minimum(1 + rng.poisson(1.5, N), 8).astype(np.int64)
    scale = max(T / 50.0, 1.0)
    base = np.minimum((rng.pareto(1.2, N) * scale).astype(np.int64), T - 1)
    slots = int(k.sum())
    cls = np.repeat(np.arange(N, dtype=np.int64), k)
    first = np.zeros(slots, dtype=bool)
    first[np.cumsum(k) - k] = True
    off = rng.integers(0, 6, slots)
    if wide_frac > 0.0:
        wide = rng.random(N) < wide_frac
        k_w = np.where(wide, rng.integers(9, wide_max + 1, N), k)
        k = k_w.astype(np.int64)
        slots = int(k.sum())
        cls = np.repeat(np.arange(N, dtype=np.int64), k)
        first = np.zeros(slots, dtype=bool)
        first[np.cumsum(k) - k] = True
        off = np.where(wide[cls], rng.integers(0, 4 * wide_max, slots), rng.integers(0, 6, slots))
    off[first] = 0
    loc = (base[cls] + off) % T
    key = np.unique(cls * T + loc)  # de-duplicate, sorted class-major then locus
    pair_class = key // T
    pair_locus = key % T
    P = key.shape[0]
    full = (1 << H) - 1
    mask = np.where(rng.random(P) < 0.5, full, rng.integers(1, full + 1, P)).astype(np.uint8)
    count = rng.geometric(0.2, N).astype(np.float64)

    data = SynthData(T=T, H=H, N=N, pair_class=pair_class, pair_locus=pair_locus, pair_mask=mask,
                     count=count, gene_of=gene_of, lengths=lengths, hname=HAPLOTYPES[:H])
    data.lname = [f"T{t:07d}" for t in range(T)]
    data.gname = [f"G{g:07d}" for g in range(int(gene_of.max()) + 1)]
    if with_genotype:
        gi = np.sort(geno_idx, axis=1)
        data.genotype = [HAPLOTYPES[a] + HAPLOTYPES[b] for a, b in gi]
    return data


def to_csc_list(d: SynthData, dtype=np.float64):
    """H scipy CSC matrices (N x T), values 1.0 on the incidence pattern."""
    from scipy.sparse import csc_matrix

    mats = []
    order = np.argsort(d.pair_locus, kind="stable")  # pairs are class-sorted, so this is (locus, class) order
    rows_all = d.pair_class[order]
    cols_all = d.pair_locus[order]
    mask_all = d.pair_mask[order]
    idx_t = np.int32 if max(d.N, d.pairs) < 2**31 - 1 else np.int64
    for h in range(d.H):
        sel = ((mask_all >> h) & 1).astype(bool)
        rows = rows_all[sel]
        indptr = np.zeros(d.T + 1, dtype=np.int64)
        indptr[1:] = np.cumsum(np.bincount(cols_all[sel], minlength=d.T))
        m = csc_matrix((np.ones(rows.size, dtype=dtype), rows.astype(idx_t), indptr.astype(idx_t)),
                       shape=(d.N, d.T))
        m.has_sorted_indices = True
        mats.append(m)
    return mats


def to_apm(d: SynthData):
    """`gbrs_b200.AlignmentPropertyMatrix` holding `d` (what loading the equivalent EMASE file would give)."""
    from .apm import AlignmentPropertyMatrix

    apm = AlignmentPropertyMatrix.from_csc(to_csc_list(d), list(d.hname), list(d.lname), count=d.count.copy())
    apm.gname = np.array(d.gname)
    apm.groups = d.groups()
    apm.num_groups = len(d.gname)
    return apm


def write_group_file(d: SynthData, path: str) -> None:
    """gene<TAB>t1<TAB>t2... (reference `AlignmentPropertyMatrix.__load_groups`,
    `src/gbrs/emase/AlignmentPropertyMatrix.py:113-130`)."""
    with open(path, "w") as fh:
        for g, tids in enumerate(d.groups()):
            fh.write(d.gname[g] + "\t" + "\t".join(d.lname[t] for t in tids) + "\n")


def write_length_file(d: SynthData, path: str) -> None:
    """locus_hap<TAB>length (reference `EMfactory.prepare`, `src/gbrs/emase/EMfactory.py:61-84`)."""
    with open(path, "w") as fh:
        if d.H > 1:
            for t in range(d.T):
                for h in range(d.H):
                    fh.write(f"{d.lname[t]}_{d.hname[h]}\t{int(d.lengths[t, h])}\n")
        else:
            for t in range(d.T):
                fh.write(f"{d.lname[t]}\t{int(d.lengths[t, 0])}\n")


def write_genotype_file(d: SynthData, path: str) -> None:
    """#Gene_ID<TAB>Diplotype (reference `quantify`, `src/gbrs/gbrs/emase_utils.py:259-269`)."""
    assert d.genotype is not None
    with open(path, "w") as fh:
        fh.write("#Gene_ID\tDiplotype\n")
        for g, gt in enumerate(d.genotype):
            fh.write(f"{d.gname[g]}\t{gt}\n")


def effective_lengths(d: SynthData, read_length: int = 100) -> np.ndarray:
    """H x T effective lengths, max(len - read_length + 1, 1) (reference `EMfactory.py:77,89`)."""
    return np.maximum(d.lengths - read_length + 1.0, 1.0).T.copy()


def genotype_mask(d: SynthData) -> np.ndarray:
    """H x T 0/1 mask of the two diplotype haplotypes of each gene (reference
    `src/gbrs/gbrs/emase_utils.py:247-269`)."""
    assert d.genotype is not None
    hid = {h: i for i, h in enumerate(d.hname)}
    gm = np.zeros((d.H, d.T))
    for g, tids in enumerate(d.groups()):
        for c in d.genotype[g]:
            gm[hid[c], tids] = 1.0
    return gm


# ----------------------------------------------------------------------------------------------------------------------
# inputs of `gbrs reconstruct` (the step after quantify: gene-level TPM -> diplotype per gene along each chromosome)
# ----------------------------------------------------------------------------------------------------------------------
@dataclass
class SynthReconstruct:
    """Logical content of the files `reconstruct` reads (reference src/gbrs/gbrs/gbrs_utils.py:382-470): chromosome
    list (`ref.fa.fai`), genes in genome order per chromosome (`ref.gene_pos.ordered.npz`), log transition matrices
    per chromosome (`tranprob.*.npz`, [steps][S][S]), alignment specificity per gene (`avecs.npz`, H x H) and the
    gene-level TPM table of one sample (`*.genes.tpm`)."""
    hname: tuple
    chroms: list  # chromosome names in fai order (may include names without genes / transition matrices)
    chrlen: dict  # name -> length (only written to the fai file)
    genes: dict  # chrom -> list of gene ids in genome order
    gpos: dict  # chrom -> list of int positions
    tprob: dict  # chrom -> float64 [steps][S][S] log transition probabilities
    avecs: dict  # gene id -> float64 [H][H]
    expr: dict  # gene id -> float64 [H] (dict order = row order of the TPM file)
    truth: dict  # chrom -> int array, the diplotype index the expression was simulated from

    @property
    def H(self) -> int:
        return len(self.hname)

    @property
    def S(self) -> int:
        return self.H * (self.H + 1) // 2


def diplotype_pairs(H: int) -> list:
    """(i, j), i <= j, in itertools.combinations_with_replacement order (reference gbrs_utils.py:452-454)."""
    return [(i, j) for i in range(H) for j in range(i, H)]


def generate_reconstruct(genes_per_chrom=(40, 25, 33), H: int = 8, sample_index: int = 0, extra_tprob_step=(),
                         frac_low: float = 0.15, frac_no_avec: float = 0.1, switch_rate: float = 0.03,
                         empty_chrom: bool = True) -> SynthReconstruct:
    """One sample.  Chromosomes are named 1, 2, ... and X (last).  `extra_tprob_step`: chromosome names whose
    transition file carries as many matrices as genes (the legacy layout the reference's back-trace special-cases,
    gbrs_utils.py:585-588) instead of genes - 1.  `frac_low` genes are expressed below any sensible threshold (null
    emission), `frac_no_avec` genes have no alignment-specificity entry (naive emission)."""
    shared = np.random.default_rng(BASE_SEED + 7919)
    rng = np.random.default_rng(BASE_SEED + 104729 + sample_index)
    hname = HAPLOTYPES[:H]
    dip = diplotype_pairs(H)
    S = len(dip)
    names = [str(i + 1) for i in range(len(genes_per_chrom) - 1)] + ["X"] if len(genes_per_chrom) > 1 else ["1"]
    chroms, chrlen, genes, gpos, tprob, avecs, expr, truth = [], {}, {}, {}, {}, {}, {}, {}
    gcount = 0
    for c, n in zip(names, genes_per_chrom):
        chroms.append(c)
        pos = np.sort(shared.integers(1000, 1000 + 40000 * max(n, 1), n))
        chrlen[c] = int(pos[-1] + 5000) if n else 5000
        genes[c] = [f"ENSG{gcount + i:08d}" for i in range(n)]
        gcount += n
        gpos[c] = [int(p) for p in pos]
        steps = n if c in extra_tprob_step else max(n - 1, 0)
        r = shared.uniform(0.002, 0.08, steps)  # probability of leaving the current diplotype
        q = shared.dirichlet(np.ones(S) * 0.5, size=(steps, S))
        p = (1.0 - r)[:, None, None] * np.eye(S)[None] + r[:, None, None] * q
        tprob[c] = np.log(p / p.sum(axis=2, keepdims=True))
        for g in genes[c]:
            if shared.random() >= frac_no_avec:
                a = np.eye(H) + shared.gamma(0.3, 0.08, (H, H))  # mostly specific, some cross-alignment
                avecs[g] = a / np.linalg.norm(a, axis=1, keepdims=True)
        # sample-specific: diplotype path and expression
        state = np.zeros(n, dtype=np.int64)
        if n:
            state[0] = rng.integers(0, S)
            for i in range(1, n):
                state[i] = rng.integers(0, S) if rng.random() < switch_rate else state[i - 1]
        truth[c] = state
        for i, g in enumerate(genes[c]):
            a = avecs.get(g, np.eye(H))
            h1, h2 = dip[state[i]]
            level = rng.lognormal(2.5, 1.2)
            if rng.random() < frac_low:
                level = rng.uniform(0.0, 0.5)
            w = rng.beta(8, 8)
            v = level * (w * a[h1] + (1 - w) * a[h2]) * rng.lognormal(0.0, 0.15, H)
            expr[g] = np.round(v, 6)
    if empty_chrom:  # a chromosome of the fai file without genes or transition matrices (e.g. Y, MT): skipped
        chroms.append("MT")
        chrlen["MT"] = 16299
    # the TPM table also lists genes that are on no chromosome of the gene-position file
    for i in range(3):
        expr[f"ENSGX{i:07d}"] = np.round(rng.gamma(1.0, 3.0, H), 6)
    return SynthReconstruct(hname=hname, chroms=chroms, chrlen=chrlen, genes=genes, gpos=gpos, tprob=tprob, avecs=avecs,
                            expr=expr, truth=truth)


def write_reconstruct_files(d: SynthReconstruct, directory: str, prefix: str = "") -> dict:
    """Write the five input files in the reference's formats; returns their paths (`data_dir` is what $GBRS_DATA must
    point at for `ref.fa.fai`)."""
    import os

    paths = {"data_dir": directory,
             "fai": os.path.join(directory, "ref.fa.fai"),
             "gpos": os.path.join(directory, prefix + "ref.gene_pos.ordered.npz"),
             "tprob": os.path.join(directory, prefix + "tranprob.npz"),
             "avecs": os.path.join(directory, prefix + "avecs.npz"),
             "expr": os.path.join(directory, prefix + "sample.genes.tpm")}
    with open(paths["fai"], "w") as fh:
        off = 0
        for c in d.chroms:
            fh.write(f"{c}\t{d.chrlen[c]}\t{off + len(c) + 2}\t60\t61\n")
            off += d.chrlen[c]
    # get_transition_prob saves lists of (gene id, position) tuples: numpy makes them [n][2] string arrays
    # (gbrs_utils.py:262-265)
    np.savez_compressed(paths["gpos"], **{c: np.array([(g, p) for g, p in zip(d.genes[c], d.gpos[c])])
                                          for c in d.genes if len(d.genes[c])})
    np.savez_compressed(paths["tprob"], **{c: d.tprob[c] for c in d.tprob if len(d.genes[c])})
    np.savez_compressed(paths["avecs"], **d.avecs)
    with open(paths["expr"], "w") as fh:
        fh.write("locus\t" + "\t".join(d.hname) + "\ttotal\n")
        for g, v in d.expr.items():
            fh.write(g + "\t" + "\t".join(repr(float(x)) for x in v) + "\t" + repr(float(v.sum())) + "\n")
    return paths
