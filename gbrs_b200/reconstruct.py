"""`reconstruct` -- genome reconstruction from gene-level TPM, the step after `quantify` (it writes the genotype table that
`quantify -G` / `stencil` read).

Same call surface, input files and output files as the reference's `reconstruct`
(/root/reference/src/gbrs/gbrs/gbrs_utils.py:382-609): per gene an emission log-probability over the H(H+1)/2 diplotypes
(null model for genes below `expr_threshold`, naive specificity for genes without an `avecs` entry, :462-488), then per
chromosome a scaled forward / backward pass, the posterior (`<outbase>.genoprobs.npz`), Viterbi scores and the
reference's back-trace (`<outbase>.genotypes.tsv`, `<outbase>.genotypes.npz`).

The reference walks the genes of every chromosome in Python with numpy operations on 36 x 36 matrices.  Here the host
only parses the files and lays the chains out (`build_plan`); emissions and chains run on the GPU through the C ABI
(`gbrs_hmm_emission`, `gbrs_hmm_run`; gbrs_b200/csrc/hmm_kernels.cu), one thread block per (sample, chromosome) --
`reconstruct_cohort` runs many samples in one launch and shares the transition matrices between them.  PyTorch tensors
are device buffers only.  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os
from collections import OrderedDict
from dataclasses import dataclass
from itertools import combinations_with_replacement

import numpy as np

from . import _lib, utils

logger = utils.get_logger("gbrs")

CHAIN_DTYPE = np.dtype([("gene0", "<i8"), ("tprob0", "<i8"), ("n_genes", "<i4"), ("n_steps", "<i4"), ("state0", "<i8")])
assert CHAIN_DTYPE.itemsize == C.sizeof(_lib.HmmChain)


def get_chromosome_info(data_dir: str) -> "OrderedDict[str, int]":
    """Chromosome names and lengths from `$GBRS_DATA/ref.fa.fai` (gbrs_utils.py:22-37)."""
    fai_file = os.path.join(data_dir, "ref.fa.fai")
    try:
        chr_lens = OrderedDict()
        with open(fai_file) as fh:
            for line in fh:
                item = line.split()
                if len(item) >= 2:
                    chr_lens[item[0]] = int(item[1])
        return chr_lens
    except FileNotFoundError:
        raise ValueError('Make sure if $GBRS_DATA is set correctly, and that "ref.fa.fai" is in that directory. '
                         f"Currently it is: {data_dir}") from None


def read_expression(expression_file: str):
    """Gene-level TPM table (`locus<TAB>A..H<TAB>total`): haplotype names from the header, one vector per gene
    (gbrs_utils.py:445-458)."""
    expr = dict()
    with open(expression_file) as fh:
        haplotypes = fh.readline().rstrip().split("\t")[1:-1]
        for curline in fh:
            item = curline.rstrip().split("\t")
            expr[item[0]] = np.array(list(map(float, item[1:-1])))
    return haplotypes, expr


def read_gene_order(gpos_file: str) -> dict:
    """chromosome -> gene ids in genome order from `ref.gene_pos.ordered.npz` (gbrs_utils.py:433-443)."""
    gene_pos = np.load(gpos_file)
    order = dict()
    for c in gene_pos.files:
        a = gene_pos[c]
        order[c] = np.array([g.decode() if isinstance(g, bytes) else str(g) for g in (a[:, 0] if a.ndim == 2 else
                                                                                    [x[0] for x in a])])
    return order


def initial_logprob(num_haps: int) -> np.ndarray:
    """Null-model log-probabilities (gbrs_utils.py:462-469)."""
    init_vec = []
    for h1, h2 in combinations_with_replacement(range(num_haps), 2):
        init_vec.append(np.log((1.0 if h1 == h2 else 2.0) / (num_haps * num_haps)))
    return np.array(init_vec)


@dataclass
class HmmPlan:
    """Host tables of one launch: S samples x C chromosomes = chains laid one after the other along the gene axis."""
    H: int
    chroms: list  # chromosome names that are processed, in fai order
    gene_ids: dict  # chrom -> np.array of gene ids
    n_samples: int
    chains: np.ndarray  # CHAIN_DTYPE [n_samples * len(chroms)], sample-major
    expr: np.ndarray  # float64 [genes_total][H]
    avec_index: np.ndarray  # int32 [genes_total]
    avecs: np.ndarray  # float64 [n_avec][H][H]
    init: np.ndarray  # float64 [S]
    tprob: np.ndarray  # float64 [matrices][S][S], the processed chromosomes' files back to back
    n_states_out: int

    @property
    def S(self) -> int:
        return self.H * (self.H + 1) // 2

    @property
    def genes_per_sample(self) -> int:
        return int(sum(len(self.gene_ids[c]) for c in self.chroms))

    def chain_of(self, sample: int, chrom_index: int):
        return self.chains[sample * len(self.chroms) + chrom_index]

    def launch_order(self) -> np.ndarray:
        """The chain records longest first: blocks are scheduled in index order and chromosomes differ several-fold in
        gene count, so the long chains must not start last.  Records carry their own offsets; the order is free."""
        return np.ascontiguousarray(self.chains[np.argsort(-self.chains["n_genes"].astype(np.int64), kind="stable")])


def build_plan(chrom_names, gene_order, tprob, avecs, expr_list, num_haps) -> HmmPlan:
    """`chrom_names`: fai order; `gene_order`: chrom -> gene ids; `tprob`: mapping chrom -> [steps][S][S] (an NpzFile or a
    dict); `avecs`: mapping gene -> [H][H]; `expr_list`: one {gene -> TPM vector} per sample.  A chromosome is processed
    if the transition file has it (gbrs_utils.py:501, :530).  Raises KeyError for a gene of a processed chromosome that
    the expression table lacks, as the reference does (`eprob[gid]`, :508)."""
    H = int(num_haps)
    if not 1 <= H <= _lib.GBRS_HPAD:
        raise NotImplementedError("1..8 haplotypes are supported by the diplotype HMM kernels")
    S = H * (H + 1) // 2
    tfiles = set(tprob.files) if hasattr(tprob, "files") else set(tprob.keys())
    afiles = set(avecs.files) if hasattr(avecs, "files") else set(avecs.keys())
    chroms = [c for c in chrom_names if c in tfiles]
    gene_ids, mats, tprob0, steps = {}, [], {}, {}
    m0 = 0
    for c in chroms:
        gene_ids[c] = np.asarray(gene_order[c])  # KeyError if the gene-position file lacks the chromosome (:504)
        t = np.ascontiguousarray(tprob[c], dtype=np.float64).reshape(-1, S, S)
        n = len(gene_ids[c])
        if n < 1:
            raise ValueError(f"chromosome {c} has no genes in the gene-position file")
        if len(t) < n - 1:
            raise IndexError(f"chromosome {c}: {n} genes need {n - 1} transition matrices, the file has {len(t)}")
        mats.append(t)
        tprob0[c], steps[c] = m0, len(t)
        m0 += len(t)
    per_sample = sum(len(gene_ids[c]) for c in chroms)
    all_ids = [g for c in chroms for g in gene_ids[c]]
    # alignment specificity: one matrix per distinct gene that has an entry, shared by all samples
    arow, amats = {}, []
    for g in all_ids:
        if g in afiles and g not in arow:
            arow[g] = len(amats)
            amats.append(np.asarray(avecs[g], dtype=np.float64).reshape(H, H))
    aidx_one = np.array([arow.get(g, -1) for g in all_ids], dtype=np.int32)
    n_samples = len(expr_list)
    expr = np.zeros((n_samples * per_sample, H))
    for s, table in enumerate(expr_list):
        base = s * per_sample
        for i, g in enumerate(all_ids):
            v = table[g]
            if len(v) != H:
                raise ValueError(f"gene {g}: {len(v)} expression values for {H} haplotypes")
            expr[base + i] = v
    chains = np.zeros(n_samples * len(chroms), dtype=CHAIN_DTYPE)
    s0 = 0
    for s in range(n_samples):
        g0 = s * per_sample
        for ci, c in enumerate(chroms):
            n = len(gene_ids[c])
            rec = chains[s * len(chroms) + ci]
            rec["gene0"], rec["tprob0"], rec["n_genes"], rec["n_steps"], rec["state0"] = g0, tprob0[c], n, steps[c], s0
            g0 += n
            s0 += min(n, steps[c]) + 1
    return HmmPlan(H=H, chroms=chroms, gene_ids=gene_ids, n_samples=n_samples, chains=chains, expr=expr,
                   avec_index=np.tile(aidx_one, n_samples), init=initial_logprob(H),
                   avecs=np.array(amats).reshape(-1, H, H) if amats else np.zeros((0, H, H)),
                   tprob=np.concatenate(mats) if mats else np.zeros((0, S, S)), n_states_out=s0)


def run_plan_on_device(plan: HmmPlan, expr_threshold: float, sigma: float, device=None, keep_work: bool = False) -> dict:
    """Upload the plan, run the emission kernel and the chain kernel, fetch posterior and states.  Returns gamma
    [genes][S], states [n_states_out] (and eprob / alpha / scaler / delta with `keep_work`)."""
    import torch

    lib = _lib.load()
    if not torch.cuda.is_available():
        raise _lib.GbrsCudaError("no CUDA device is available; the gbrs_b200 genome reconstruction has no CPU fallback")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    S, G = plan.S, plan.expr.shape[0]
    f64 = torch.float64

    def up(a):
        a = np.ascontiguousarray(a)
        if a.size == 0:  # keep a valid pointer
            return torch.zeros(16, dtype=torch.uint8, device=dev)
        return torch.from_numpy(a.view(np.uint8).reshape(-1)).to(dev)

    with torch.cuda.device(dev):
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        d_expr, d_aidx, d_avecs, d_init = up(plan.expr), up(plan.avec_index), up(plan.avecs), up(plan.init)
        d_tprob, d_chains = up(plan.tprob), up(plan.launch_order())
        eprob = torch.empty((max(G, 1), S), dtype=f64, device=dev)
        alpha = torch.empty((max(G, 1), S), dtype=f64, device=dev)
        gamma = torch.empty((max(G, 1), S), dtype=f64, device=dev)
        delta = torch.empty((max(G, 1), S), dtype=f64, device=dev)
        scaler = torch.empty(max(G, 1), dtype=f64, device=dev)
        n_mat = int(plan.tprob.shape[0])
        tlin = torch.empty(max(n_mat * S * S, 1), dtype=f64, device=dev)
        backptr = torch.zeros((max(G, 1), S), dtype=torch.uint8, device=dev)
        states = torch.zeros(max(plan.n_states_out, 1), dtype=torch.int32, device=dev)
        _lib.check(lib.gbrs_hmm_emission(G, plan.H, d_expr.data_ptr(), d_avecs.data_ptr(), d_aidx.data_ptr(),
                                         d_init.data_ptr(), float(expr_threshold), float(sigma), eprob.data_ptr(),
                                         stream))
        _lib.check(lib.gbrs_hmm_run(len(plan.chains), d_chains.data_ptr(), plan.H, d_init.data_ptr(), eprob.data_ptr(),
                                    d_tprob.data_ptr(), n_mat, tlin.data_ptr(), alpha.data_ptr(), scaler.data_ptr(),
                                    gamma.data_ptr(), delta.data_ptr(), backptr.data_ptr(), states.data_ptr(), stream))
        out = {"gamma": gamma[:G].cpu().numpy(), "states": states[: plan.n_states_out].cpu().numpy()}
        if keep_work:
            out.update(eprob=eprob[:G].cpu().numpy(), alpha=alpha[:G].cpu().numpy(), scaler=scaler[:G].cpu().numpy(),
                       delta=delta[:G].cpu().numpy())
    return out


def collect_sample(plan: HmmPlan, result: dict, sample: int, genotypes) -> tuple:
    """Per-chromosome outputs of one sample in the reference's shapes: gamma[c] = S x genes (:552-558),
    viterbi_states[c] = list of diplotype names (:578-596), gtcall_g = gene -> diplotype of the called genes."""
    gamma, viterbi_states, gtcall_g = dict(), dict(), dict()
    for ci, c in enumerate(plan.chroms):
        ch = plan.chain_of(sample, ci)
        g0, n = int(ch["gene0"]), int(ch["n_genes"])
        called = min(n, int(ch["n_steps"]))
        st = result["states"][int(ch["state0"]): int(ch["state0"]) + called + 1]
        gamma[c] = np.ascontiguousarray(result["gamma"][g0: g0 + n].T)
        viterbi_states[c] = [genotypes[s] for s in st]
        for i in range(called):
            gtcall_g[plan.gene_ids[c][i]] = genotypes[st[i]]
    return gamma, viterbi_states, gtcall_g


def write_outputs(outbase, gamma, viterbi_states, gtcall_g) -> None:
    """File names and formats of gbrs_utils.py:399-406, :560-561, :598-607."""
    if outbase is None:
        out_gtype = "gbrs.reconstructed.genotypes.tsv"
        out_gprob = "gbrs.reconstructed.genoprobs.npz"
    else:
        out_gtype = f"{outbase}.genotypes.tsv"
        out_gprob = f"{outbase}.genoprobs.npz"
    out_gtype_ordered = f"{os.path.splitext(out_gtype)[0]}.npz"
    logger.info(f"Saving Reconstructed Genotype Probabilities: {out_gprob}")
    np.savez_compressed(out_gprob, **gamma)
    logger.info(f"Saving Reconstructed Genotypes: {out_gtype}")
    with open(out_gtype, "w") as fhout:
        fhout.write("#Gene_ID\tDiplotype\n")
        for g in sorted(gtcall_g.keys()):
            fhout.write(f"{g}\t{gtcall_g[g]}\n")
    logger.info(f"Saving Reconstructed Ordered Genotypes: {out_gtype_ordered}")
    np.savez_compressed(out_gtype_ordered, **viterbi_states)


def reconstruct_cohort(expression_files, tprob_file: str, avec_file: str = None, gpos_file: str = None,
                       expr_threshold: float = 1.5, sigma: float = 0.12, outbases=None, device=None,
                       group=None) -> None:
    """`reconstruct` for many samples that share the reference files: one emission launch and one chain launch for
    all (sample, chromosome) chains.  With a torch.distributed process `group` (or an initialised default group) the
    samples are dealt round-robin over the ranks -- replicas only, no collective: chains are independent."""
    data_dir = os.getenv("GBRS_DATA", ".")
    expression_files = list(expression_files)
    if outbases is None:
        outbases = [None] if len(expression_files) == 1 else [os.path.splitext(f)[0] for f in expression_files]
    if len(outbases) != len(expression_files):
        raise ValueError("one outbase per expression file is required")
    rank, world = 0, 1
    try:
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized():
            rank, world = dist.get_rank(group), dist.get_world_size(group)
    except ImportError:
        pass
    mine = list(range(rank, len(expression_files), world))
    if avec_file is None:
        avec_file = os.path.join(data_dir, "avecs.npz")
    if gpos_file is None:
        gpos_file = os.path.join(data_dir, "ref.gene_pos.ordered.npz")
    logger.info(f"Expression File(s): {', '.join(expression_files[i] for i in mine)}")
    logger.info(f"Transition Probabilities File: {tprob_file}")
    logger.info(f"Alignment Specificity File: {avec_file}")
    logger.info(f"Gene Position File: {gpos_file}")
    logger.info(f"Expression Threshold: {expr_threshold}")
    logger.info(f"Sigma: {sigma}")
    logger.info("Loading chromosome information")
    chrlens = get_chromosome_info(data_dir)
    logger.info(f"Loading gene meta data: {gpos_file}")
    gene_order = read_gene_order(gpos_file)
    haplotypes, tables = None, []
    for i in mine:
        logger.info(f"Loading expression level data: {expression_files[i]}")
        h, expr = read_expression(expression_files[i])
        if haplotypes is not None and h != haplotypes:
            raise ValueError(f"{expression_files[i]}: haplotype columns differ from the first sample's")
        haplotypes = h
        tables.append(expr)
    if not tables:
        return
    genotypes = [h1 + h2 for h1, h2 in combinations_with_replacement(haplotypes, 2)]
    # the chromosomes of the fai file that the transition file has (gbrs_utils.py:501), then only the alignment
    # specificity of their genes: members are read in bulk (utils.read_npz_members) instead of one np.load item at a time
    logger.info(f"Loading transition probabilities: {tprob_file}")
    wanted_chroms = set(chrlens.keys())
    tprob = utils.read_npz_members(tprob_file, lambda name: name in wanted_chroms)
    logger.info(f"Loading alignment specificity: {avec_file}")
    wanted_genes = set(g for c in tprob for g in gene_order.get(c, ()))
    avecs = utils.read_npz_members(avec_file, lambda name: name in wanted_genes)
    plan = build_plan(list(chrlens.keys()), gene_order, tprob, avecs, tables, len(haplotypes))
    logger.info("Getting forward-backward probability and running Viterbi on the GPU")
    result = run_plan_on_device(plan, expr_threshold, sigma, device=device)
    for s, i in enumerate(mine):
        write_outputs(outbases[i], *collect_sample(plan, result, s, genotypes))
    logger.info("Done")


def reconstruct(expression_file: str, tprob_file: str, avec_file: str = None, gpos_file: str = None,
                expr_threshold: float = 1.5, sigma: float = 0.12, outbase: str = None) -> None:
    """Reconstruct the genome based upon gene-level TPM quantities (gbrs_utils.py:382-609)."""
    reconstruct_cohort([expression_file], tprob_file, avec_file=avec_file, gpos_file=gpos_file,
                       expr_threshold=expr_threshold, sigma=sigma, outbases=[outbase])
