"""Not a test: times the CPU restatement of the reference's `reconstruct` inner loops (oracle/reconstruct_oracle.py:
per-gene emissions in Python, per-gene numpy steps of the forward / backward / Viterbi passes) on one chromosome of one
synthetic sample -- the CPU baseline quoted next to tools/bench_reconstruct.py.  One JSON line.

    python -m tests.time_reconstruct_oracle [--genes 1900]"""
from __future__ import annotations

import argparse
import json
import os
import time

import numpy as np

from gbrs_b200 import synth
from oracle import reconstruct_oracle as ro


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genes", type=int, default=1900)
    args = ap.parse_args()
    d = synth.generate_reconstruct(genes_per_chrom=(args.genes, 2), H=8, empty_chrom=False)
    init = ro.initial_logprob(8)
    t0 = time.perf_counter()
    e = np.array([ro.emission_logprob(np.asarray(d.expr[g]), d.avecs.get(g), init) for g in d.genes["1"]])
    t_em = time.perf_counter() - t0
    t0 = time.perf_counter()
    ro.reconstruct_chain(init, e, d.tprob["1"])
    t_ch = time.perf_counter() - t0
    print(json.dumps({"metric": "reconstruct_gene_steps_per_s", "kind": "port", "cores": 1, "host_cores": os.cpu_count(),
                      "sample": f"one chromosome of {args.genes} genes, one sample, H = 8",
                      "value": args.genes / (t_em + t_ch), "emission_seconds": t_em, "chain_seconds": t_ch}))


if __name__ == "__main__":
    main()
