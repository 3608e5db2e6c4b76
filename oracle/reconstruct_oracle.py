"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's `reconstruct`
(/root/reference/src/gbrs/gbrs/gbrs_utils.py:382-609): per-gene emission log-probabilities over the H(H+1)/2 diplotypes,
scaled forward / backward passes and the posterior per chromosome, Viterbi scores and the reference's back-trace.

Nothing under gbrs_b200/ imports this file; tests and tools use it as the checker for the CUDA path
(gbrs_b200/csrc/hmm_kernels.cu).  Parity pin: `python -m oracle.make_golden_reconstruct` runs the UNMODIFIED reference
`reconstruct` in the build container (matplotlib, which the reference module imports at the top but does not use on this
path, is stubbed) and stores its output files in tests/golden/reconstruct_*.npz; tests/test_reconstruct_oracle.py checks
this restatement against them (posterior to 1e-12, Viterbi path and genotype table exactly).

Every function follows the reference's operation order (python `sum` = left-to-right, numpy broadcasting as written), so
on the same numpy the forward / backward values are bit-identical to the reference's.
"""
from __future__ import annotations

from itertools import combinations_with_replacement

import numpy as np

TINY = np.nextafter(0, 1)  # the reference adds this before every log of a probability (gbrs_utils.py:481, :487, :516)


def unit_vector(vector):
    """gbrs_utils.py:63-67."""
    if sum(vector) > 1e-6:
        return vector / np.linalg.norm(vector)
    return vector


def genotype_probability(aln_profile, aln_specificity, sigma=0.12):
    """gbrs_utils.py:80-100: squared distance between the unit expression vector and the unit specificity vector of
    every diplotype (homozygote: the founder's row; heterozygote: unit(v1 + v2)), gaussian kernel, normalised."""
    num_haps = len(aln_profile)
    aln_vec = unit_vector(aln_profile)
    genoprob = []
    for i in range(num_haps):
        v1 = unit_vector(aln_specificity[i])
        for j in range(i, num_haps):
            if j == i:
                genoprob.append(sum(np.power(aln_vec - v1, 2)))
            else:
                v2 = unit_vector(aln_specificity[j])
                geno_vec = unit_vector(v1 + v2)
                genoprob.append(sum(np.power(aln_vec - geno_vec, 2)))
    genoprob = np.exp(np.array(genoprob) / (-2 * sigma * sigma))
    return np.array(genoprob / sum(genoprob))


def initial_logprob(num_haps):
    """Null-model log-probabilities (gbrs_utils.py:462-469): 1/H^2 for homozygotes, 2/H^2 for heterozygotes."""
    init_vec = []
    for h1, h2 in combinations_with_replacement(range(num_haps), 2):
        if h1 == h2:
            init_vec.append(np.log(1.0 / (num_haps * num_haps)))
        else:
            init_vec.append(np.log(2.0 / (num_haps * num_haps)))
    return np.array(init_vec)


def emission_logprob(evec, avec, init_vec, expr_threshold=1.5, sigma=0.12):
    """gbrs_utils.py:472-488.  `avec` None = the gene has no alignment-specificity entry (naive vectors, sigma 0.45)."""
    num_haps = len(evec)
    if sum(evec) < expr_threshold:
        return init_vec
    if avec is None:
        naiv = np.eye(num_haps) + (np.ones((num_haps, num_haps)) - np.eye(num_haps)) * 0.0001
        return np.log(genotype_probability(evec, naiv, sigma=0.450) + TINY)
    return np.log(genotype_probability(evec, avec, sigma=sigma) + TINY)


def forward(init_vec, eprob, tprob):
    """gbrs_utils.py:498-523.  eprob [n][S] (gene-major), tprob [steps][S][S].  Returns alpha [S][n], scaler [n]."""
    n, S = eprob.shape
    alpha = np.zeros((S, n))
    scaler = np.zeros(n)
    alpha[:, 0] = init_vec + eprob[0]
    normalizer = np.log(sum(np.exp(alpha[:, 0])))
    alpha[:, 0] -= normalizer
    scaler[0] = -normalizer
    for i in range(1, n):
        alpha[:, i] = np.log(np.exp(alpha[:, i - 1] + tprob[i - 1]).sum(axis=1) + TINY) + eprob[i]
        normalizer = np.log(sum(np.exp(alpha[:, i])))
        alpha[:, i] -= normalizer
        scaler[i] = -normalizer
    return alpha, scaler


def backward(eprob, tprob, scaler):
    """gbrs_utils.py:527-548."""
    n, S = eprob.shape
    beta = np.zeros((S, n))
    beta[:, -1] = scaler[-1]
    for i in range(n - 2, -1, -1):
        beta[:, i] = np.log(np.exp(tprob[i].transpose() + beta[:, i + 1] + eprob[i + 1] + scaler[i]).sum(axis=1))
    return beta


def posterior(alpha, beta):
    """gbrs_utils.py:552-558."""
    gamma = np.exp(alpha + beta)
    return gamma / gamma.sum(axis=0)


def viterbi(init_vec, eprob, tprob):
    """gbrs_utils.py:565-596.  Returns delta [S][n] and the state list in the reference's order: the states of genes
    0 .. m-1 (m = min(n, steps): the back-trace only visits genes that have a transition matrix of their own index)
    followed by the arg-max state of the last gene.  `called` = m, the number of genes that get a genotype call."""
    n, S = eprob.shape
    delta = np.zeros((S, n))
    delta[:, 0] = init_vec + eprob[0]
    for i in range(1, n):
        delta[:, i] = (delta[:, i - 1] + tprob[i - 1]).max(axis=1) + eprob[i]
    sid = delta[:, n - 1].argmax()
    states = [int(sid)]
    m = n
    if m > len(tprob):
        m = len(tprob)
    for i in reversed(range(m)):
        sid = (delta[:, i] + tprob[i][sid]).argmax()
        states.append(int(sid))
    states.reverse()
    return delta, np.array(states, dtype=np.int64), m


def reconstruct_chain(init_vec, eprob, tprob):
    alpha, scaler = forward(init_vec, eprob, tprob)
    beta = backward(eprob, tprob, scaler)
    gamma = posterior(alpha, beta)
    delta, states, called = viterbi(init_vec, eprob, tprob)
    return {"alpha": alpha, "scaler": scaler, "beta": beta, "gamma": gamma, "delta": delta, "states": states,
            "called": called}


def reconstruct_tables(chroms, genes, tprob, avecs, expr, haplotypes, expr_threshold=1.5, sigma=0.12, eprob=None):
    """The whole of gbrs_utils.py:450-603 on in-memory tables.  `chroms`: chromosome names in fai order; `genes`:
    chrom -> gene ids in genome order; `tprob`: chrom -> [steps][S][S]; `avecs`: gene -> [H][H]; `expr`: gene -> [H].
    `eprob` (gene -> [S]) replaces the computed emissions when given (used to check the chain kernels on the device's
    own emission values).  Returns gamma / viterbi state names per chromosome and the gene -> diplotype calls."""
    H = len(haplotypes)
    genotypes = [h1 + h2 for h1, h2 in combinations_with_replacement(haplotypes, 2)]
    init_vec = initial_logprob(H)
    if eprob is None:
        eprob = {g: emission_logprob(np.asarray(v, dtype=float), avecs.get(g), init_vec, expr_threshold, sigma)
                 for g, v in expr.items()}
    gamma, vit, gtcall, detail = {}, {}, {}, {}
    for c in chroms:
        if c not in tprob:
            continue
        ids = genes[c]
        e = np.array([eprob[g] for g in ids])
        r = reconstruct_chain(init_vec, e, np.asarray(tprob[c]))
        gamma[c] = r["gamma"]
        vit[c] = [genotypes[s] for s in r["states"]]
        for i in range(r["called"]):
            gtcall[ids[i]] = genotypes[r["states"][i]]
        detail[c] = r
    return {"gamma": gamma, "viterbi": vit, "gtcall": gtcall, "eprob": eprob, "detail": detail, "genotypes": genotypes}
