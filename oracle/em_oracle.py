"""TEST INFRASTRUCTURE ONLY -- CPU (numpy) restatement of the reference's multiway EM quantifier.

This file is the parity oracle for the CUDA path.  It may be imported only by `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of `bench.py`.  The product
package `gbrs_b200` never imports it and has no CPU fallback.

PARITY PIN.  The reference ships no golden vectors or known-answer tests for this path
(`/root/reference/tests/test_gbrs.py:15-24` is an empty placeholder), so the oracle is pinned against
*the reference itself executed in the build container*: `oracle/make_golden.py` runs the unmodified
reference `EMfactory` (model 4 as-is; models 1-3 through the one-operation sparse-divide shim of
`oracle/ref_harness.py`, because the reference crashes there on current scipy -- SURVEY.md fact 3) and
commits its outputs under `tests/golden/`; `tests/test_oracle_golden.py` checks this restatement
against those vectors (<= 1e-12 relative, identical iteration counts).

The restatement keeps the reference's sequence of operations per model but works on one flat entry
list instead of H scipy CSC matrices:

    entry e = (class ent_n[e], locus ent_t[e], haplotype ent_h[e]),  value val[e]

ordered haplotype-major, then locus, then class -- i.e. the concatenation of the reference's
`data[h].indices` (`src/gbrs/emase/Sparse3DMatrix.py:26-66`).  Reference citations are given per
function; all paths are relative to /root/reference.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass
class OracleAPM:
    """Incidence pattern + class counts (reference `AlignmentPropertyMatrix`,
    src/gbrs/emase/AlignmentPropertyMatrix.py:25-111)."""
    T: int
    H: int
    N: int
    ent_n: np.ndarray
    ent_t: np.ndarray
    ent_h: np.ndarray
    count: np.ndarray | None  # None => all ones (AlignmentPropertyMatrix.py:291-293)

    @property
    def nnz(self) -> int:
        return int(self.ent_n.size)

    def weights(self) -> np.ndarray:
        return np.ones(self.N) if self.count is None else self.count


def apm_from_csc_list(mats, count) -> OracleAPM:
    """From the reference's storage: list of H CSC matrices N x T."""
    H = len(mats)
    N, T = mats[0].shape
    en, et, eh = [], [], []
    for h, m in enumerate(mats):
        m = m.tocsc()
        en.append(np.asarray(m.indices, dtype=np.int64))
        et.append(np.repeat(np.arange(T, dtype=np.int64), np.diff(m.indptr)))
        eh.append(np.full(m.indices.size, h, dtype=np.int64))
    return OracleAPM(T, H, N, np.concatenate(en), np.concatenate(et), np.concatenate(eh),
                     None if count is None else np.asarray(count, dtype=np.float64))


def apm_from_pairs(T, H, N, pair_class, pair_locus, pair_mask, count) -> OracleAPM:
    en, et, eh = [], [], []
    for h in range(H):
        sel = ((pair_mask >> h) & 1).astype(bool)
        n, t = pair_class[sel], pair_locus[sel]
        o = np.lexsort((n, t))
        en.append(n[o].astype(np.int64))
        et.append(t[o].astype(np.int64))
        eh.append(np.full(o.size, h, dtype=np.int64))
    return OracleAPM(T, H, N, np.concatenate(en), np.concatenate(et), np.concatenate(eh),
                     None if count is None else np.asarray(count, dtype=np.float64))


def apply_genotype_mask(apm: OracleAPM, gtmask: np.ndarray) -> OracleAPM:
    """`aln_mat.multiply(gtmask, axis=2)` + `eliminate_zeros()` -- src/gbrs/gbrs/emase_utils.py:271-273."""
    keep = gtmask[apm.ent_h, apm.ent_t] != 0
    return OracleAPM(apm.T, apm.H, apm.N, apm.ent_n[keep], apm.ent_t[keep], apm.ent_h[keep], apm.count)


def gene_index(T: int, groups) -> np.ndarray:
    """gene id per locus as implied by `t2t_mat` (identity plus same-gene pairs,
    src/gbrs/emase/EMfactory.py:48-59): ungrouped loci become singleton genes."""
    g = np.full(T, -1, dtype=np.int64)
    for gi, tids in enumerate(groups or []):
        g[np.asarray(tids, dtype=np.int64)] = gi
    free = np.flatnonzero(g < 0)
    g[free] = (len(groups) if groups else 0) + np.arange(free.size)
    return g


class _Keys:
    """Pre-computed segment ids for the per-class normalisers."""

    def __init__(self, apm: OracleAPM, gene_of: np.ndarray | None):
        self.nt, self.n_nt = self._ids(apm.ent_n * apm.T + apm.ent_t)
        if gene_of is not None:
            Gx = int(gene_of.max()) + 1
            g = gene_of[apm.ent_t]
            self.ng, self.n_ng = self._ids(apm.ent_n * Gx + g)
            self.ngh, self.n_ngh = self._ids((apm.ent_n * Gx + g) * apm.H + apm.ent_h)

    @staticmethod
    def _ids(key):
        _, inv = np.unique(key, return_inverse=True)
        return inv, int(inv.max()) + 1 if inv.size else 0


def _div_on_pattern(val, den):
    """element-wise division on the numerator's pattern after `eliminate_zeros()`
    (src/gbrs/emase/AlignmentPropertyMatrix.py:323-327, 350-352, 364-366): structural zeros stay 0."""
    out = np.zeros_like(val)
    nz = val != 0
    out[nz] = val[nz] / den[nz]
    return out


def sum_read(apm: OracleAPM, val: np.ndarray) -> np.ndarray:
    """`APM.sum(axis=READ)`: count-weighted column reduce -> H x T
    (src/gbrs/emase/AlignmentPropertyMatrix.py:288-298)."""
    w = val if apm.count is None else val * apm.count[apm.ent_n]
    return np.bincount(apm.ent_h * apm.T + apm.ent_t, weights=w, minlength=apm.H * apm.T).reshape(apm.H, apm.T)


def normalize_read(apm: OracleAPM, val: np.ndarray) -> np.ndarray:
    """`normalize_reads(axis=READ)` (AlignmentPropertyMatrix.py:335-342): divide by the per-class total."""
    tot = np.bincount(apm.ent_n, weights=val, minlength=apm.N)
    return val / tot[apm.ent_n]


def e_step(apm: OracleAPM, theta: np.ndarray, model: int, gene_of=None, keys: _Keys | None = None) -> np.ndarray:
    """`EMfactory.update_probability_at_read_level` (src/gbrs/emase/EMfactory.py:146-212).
    Returns the posterior value of every entry."""
    n, t, h = apm.ent_n, apm.ent_t, apm.ent_h
    val = theta[h, t].astype(np.float64)  # reset() then multiply(theta, axis=READ)  :159, Sparse3DMatrix.py:354-362
    if model == 4:  # :204-208
        return normalize_read(apm, val)
    if model not in (1, 2, 3):
        raise RuntimeError("The read normalization model should be 1, 2, 3, or 4.")
    if keys is None:
        keys = _Keys(apm, gene_of)
    # (theta * t2t): per-haplotype gene totals broadcast back to loci -> H x T
    Gx = int(gene_of.max()) + 1
    hg_gene = np.zeros((apm.H, Gx))
    for hh in range(apm.H):
        hg_gene[hh] = np.bincount(gene_of, weights=theta[hh], minlength=Gx)
    hg = hg_gene[:, gene_of]  # haplogroup_sum_mat  :167
    gamma = hg.sum(axis=0)  # (theta * t2t).sum(axis=0)  :173, :188, :200
    if model == 3:  # :192-203
        den = np.bincount(keys.ng, weights=val, minlength=keys.n_ng)  # GROUP  AlignmentPropertyMatrix.py:343-352
        val = _div_on_pattern(val, den[keys.ng])
        val = val * gamma[t]
        return normalize_read(apm, val)
    if model == 2:  # :176-191
        den = np.bincount(keys.nt, weights=val, minlength=keys.n_nt)  # LOCUS  AlignmentPropertyMatrix.py:316-327
        val = _div_on_pattern(val, den[keys.nt])
        val = val * theta.sum(axis=0)[t]
        den = np.bincount(keys.ng, weights=val, minlength=keys.n_ng)
        val = _div_on_pattern(val, den[keys.ng])
        val = val * gamma[t]
        return normalize_read(apm, val)
    # model 1  :160-175
    den = np.bincount(keys.ngh, weights=val, minlength=keys.n_ngh)  # HAPLOGROUP  AlignmentPropertyMatrix.py:353-366
    val = _div_on_pattern(val, den[keys.ngh])
    val = val * hg[h, t]
    den = np.bincount(keys.ng, weights=val, minlength=keys.n_ng)
    val = _div_on_pattern(val, den[keys.ng])
    val = val * gamma[t]
    return normalize_read(apm, val)


def prepare(apm: OracleAPM, efflen: np.ndarray | None, pseudocount: float = 0.0) -> np.ndarray:
    """`EMfactory.prepare` numeric part (src/gbrs/emase/EMfactory.py:95-111): theta0 (H x T)."""
    val = normalize_read(apm, np.ones(apm.nnz))
    theta = sum_read(apm, val)
    if efflen is not None:
        theta = np.divide(theta, efflen)
    if pseudocount > 0.0:
        s = theta.sum()
        nzloci = np.nonzero(theta)[1]
        theta[:, nzloci] += pseudocount
        theta *= s / theta.sum()
    return theta


def effective_length_table(lengths_TH: np.ndarray, read_length: int = 100) -> np.ndarray:
    """max(len - read_length + 1, 1), transposed to H x T (EMfactory.py:77, :89); all entries must be
    positive (:91-94)."""
    eff = np.maximum(np.asarray(lengths_TH, dtype=np.float64) - read_length + 1.0, 1.0).T.copy()
    if not np.all(eff > 0.0):
        raise RuntimeError("There exist transcripts missing length information.")
    return eff


def run(apm: OracleAPM, theta0: np.ndarray, model: int, efflen: np.ndarray | None, gene_of=None,
        tol: float = 0.001, max_iters: int = 999) -> dict:
    """`EMfactory.run` (src/gbrs/emase/EMfactory.py:234-287) with `update_allelic_expression`
    (:214-232) inlined.  Returns theta, iteration count, exact err_sum per iteration and the
    expected read counts of the last posterior (`report_read_counts`, :302)."""
    theta = np.array(theta0, dtype=np.float64, copy=True)
    keys = _Keys(apm, gene_of) if model in (1, 2, 3) else None
    errs = []
    counts = None
    num_iters = 0
    err_sum = 1000000.0
    target_err = 1000000.0 * tol
    with np.errstate(all="raise", under="ignore"):  # :256-257
        while err_sum > target_err and num_iters < max_iters:
            prev = theta.sum(axis=0)
            prev *= 1000000.0 / prev.sum()
            val = e_step(apm, theta, model, gene_of, keys)
            counts = sum_read(apm, val)
            theta = np.divide(counts, efflen) if efflen is not None else counts.copy()
            curr = theta.sum(axis=0)
            curr *= 1000000.0 / curr.sum()
            err_sum = float(np.abs(curr - prev).sum())
            errs.append(err_sum)
            num_iters += 1
    if counts is None:  # max_iters == 0: report_read_counts would sum the prepare() posterior
        counts = sum_read(apm, normalize_read(apm, np.ones(apm.nnz)))
    return dict(theta=theta, iters=num_iters, errs=np.array(errs), counts=counts)


def tpm(theta: np.ndarray) -> np.ndarray:
    """`report_depths(tpm=True)` scaling (EMfactory.py:352-354)."""
    return theta * (1000000.0 / theta.sum())


def group_sum(mat_HT: np.ndarray, groups, G: int) -> np.ndarray:
    """`X * grp_conv_mat` (EMfactory.py:305, :349): H x G; loci in no group vanish."""
    out = np.zeros((mat_HT.shape[0], G))
    for gi, tids in enumerate(groups):
        out[:, gi] = mat_HT[:, tids].sum(axis=1)
    return out


def gene_tpm_after_isoform_report(theta: np.ndarray, groups, G: int) -> np.ndarray:
    """Gene-level TPM as `quantify` produces it: the isoform report has already TPM-scaled theta in
    place (EMfactory.py:352-354 aliasing), then the gene table is `theta_tpm * grp_conv_mat`
    rescaled to 1e6 again (:349, :354)."""
    g = group_sum(tpm(theta), groups, G)
    return g * (1000000.0 / g.sum())


def alignment_counts(apm: OracleAPM) -> dict:
    """`report_alignment_counts` columns (src/gbrs/emase/AlignmentPropertyMatrix.py:389-459):
    aln (H x T), uniq (H x T; classes with exactly one alignment overall), locus_uniq (T; classes
    hitting exactly one locus, any haplotypes)."""
    c = apm.weights()
    ones = np.ones(apm.nnz)
    aln = sum_read(apm, ones)
    per_class = np.bincount(apm.ent_n, minlength=apm.N)
    u = per_class[apm.ent_n] == 1
    uniq = np.bincount(apm.ent_h[u] * apm.T + apm.ent_t[u], weights=c[apm.ent_n[u]],
                       minlength=apm.H * apm.T).reshape(apm.H, apm.T)
    nt = np.unique(apm.ent_n * apm.T + apm.ent_t)
    n_of, t_of = nt // apm.T, nt % apm.T
    loci_per_class = np.bincount(n_of, minlength=apm.N)
    lu = loci_per_class[n_of] == 1
    locus_uniq = np.bincount(t_of[lu], weights=c[n_of[lu]], minlength=apm.T)
    return dict(aln=aln, uniq=uniq, locus_uniq=locus_uniq)


def bundle(apm: OracleAPM, groups) -> OracleAPM:
    """`_bundle_inline(reset=True)` (AlignmentPropertyMatrix.py:155-188): loci -> genes, values reset to
    1 on the bundled pattern; loci in no group disappear."""
    g = np.full(apm.T, -1, dtype=np.int64)
    for gi, tids in enumerate(groups):
        g[np.asarray(tids, dtype=np.int64)] = gi
    G = len(groups)
    ge = g[apm.ent_t]
    keep = ge >= 0
    key = np.unique((apm.ent_h[keep] * G + ge[keep]) * apm.N + apm.ent_n[keep])
    return OracleAPM(G, apm.H, apm.N, key % apm.N, (key // apm.N) % G, key // (apm.N * G), apm.count)
