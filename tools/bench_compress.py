"""Throughput of the GPU equivalence-class construction (`gbrs compress`, gbrs_b200/compress.py) at scale, next to the
CPU restatement of the reference's per-read loop (oracle/compress_oracle.py) on a bounded sample.  One JSON line.

    python tools/bench_compress.py [--classes 1000000] [--loci 80000]

Reads are made by replicating the classes of the canonical synthetic generator `count` times and shuffling, so the
answer is known: the classes come back with their counts (checked)."""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--classes", type=int, default=1_000_000)
    ap.add_argument("--loci", type=int, default=80_000)
    ap.add_argument("--cpu-sample", type=int, default=200_000)
    args = ap.parse_args()
    import torch

    from gbrs_b200 import compress as cz
    from gbrs_b200 import synth

    d = synth.generate(T=args.loci, N=args.classes, H=8)
    k = np.bincount(d.pair_class, minlength=d.N)
    class_ptr = np.concatenate(([0], np.cumsum(k)))
    class_words = (d.pair_locus.astype(np.uint32) | (d.pair_mask.astype(np.uint32) << np.uint32(24)))
    rng = np.random.default_rng(7)
    read_class = np.repeat(np.arange(d.N), d.count.astype(np.int64))
    rng.shuffle(read_class)
    n = read_class.shape[0]
    lens = k[read_class]
    rowptr = np.concatenate(([0], np.cumsum(lens))).astype(np.int64)
    off = np.arange(int(lens.sum()), dtype=np.int64) - np.repeat(rowptr[:-1], lens)
    words = class_words[np.repeat(class_ptr[:-1][read_class], lens) + off]

    torch.cuda.synchronize()
    cz.equivalence_classes(rowptr[:1001], words[: rowptr[1000]], None)  # warm-up: context, module load
    t0 = time.perf_counter()
    cls, first, ccount = cz.equivalence_classes(rowptr, words, None)
    t_gpu = time.perf_counter() - t0
    # the generator may draw the same pattern for several of its classes: those merge.  Expected answer from the
    # class list: distinct rows, their summed counts, numbered by the first read that shows them.
    rows = {}
    for c in range(d.N):
        rows.setdefault(class_words[class_ptr[c]:class_ptr[c + 1]].tobytes(), []).append(c)
    group_of_class = np.empty(d.N, dtype=np.int64)
    group_count = np.zeros(len(rows))
    for g, members in enumerate(rows.values()):
        group_of_class[members] = g
        group_count[g] = d.count[members].sum()
    read_group = group_of_class[read_class]
    assert len(first) == len(rows) and ccount.sum() == n
    assert np.all(np.diff(first.astype(np.int64)) > 0) and np.array_equal(cls[first], np.arange(len(first)))
    assert np.array_equal(read_group[first][cls], read_group)            # membership
    assert np.array_equal(ccount, group_count[read_group[first]])        # counts
    seen_first = np.full(len(rows), n, dtype=np.int64)
    np.minimum.at(seen_first, read_group, np.arange(n))
    assert np.array_equal(np.sort(seen_first), first.astype(np.int64))   # representatives = first appearances

    # CPU: the reference's loop (string key per read, dict) restated, on the first `cpu_sample` reads
    m = min(args.cpu_sample, n)
    t0 = time.perf_counter()
    ec = {}
    for r in range(m):
        key = words[rowptr[r]:rowptr[r + 1]].tobytes()
        ec[key] = ec.get(key, 0.0) + 1.0
    t_cpu = time.perf_counter() - t0
    line = {"metric": "compress_reads_per_s", "unit": "reads/s", "reads": int(n), "classes": int(d.N),
            "pair_words": int(words.shape[0]), "gpu": {"value": n / t_gpu, "seconds": t_gpu,
            "what": "H2D of the read rows + hash / radix sort / exact grouping / numbering + D2H, through gbrs_ec_build"},
            "cpu_baseline": {"value": m / t_cpu, "seconds": t_cpu, "cores": 1, "kind": "port",
                             "sample": f"first {m} reads, dict keyed by the row bytes (cheaper than the reference's "
                                       "string keys, emase_utils.py:62-72)"}}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
