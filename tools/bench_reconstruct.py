"""Throughput of the GPU diplotype HMM (`gbrs reconstruct`, gbrs_b200/reconstruct.py) at mouse scale -- 20 chromosomes,
~25k genes, 36 diplotypes -- for one sample and for a cohort sharing the transition matrices.  One JSON line.  (CPU figure of the
reference's per-gene numpy loop: `python -m tests.time_reconstruct_oracle`.)

    python tools/bench_reconstruct.py [--samples 96] [--genes 25000] [--repeat 5]

Device time = CUDA events on the launch stream around the emission kernel + the chain kernel, tables resident;
`e2e` adds the upload of expression / transition tables and the download of posterior and states.  Algorithmic bytes of
the chain kernel: every transition matrix three times (forward, backward, Viterbi) per sample + emission, forward, posterior, score and
back-pointer tables once each; in a cohort the matrices are shared, so DRAM traffic should stay near one copy of them
(they fit the 126 MB L2 only per chromosome, not as a whole: 25k x 10 KB = 259 MB)."""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", type=int, default=96)
    ap.add_argument("--genes", type=int, default=25000)
    ap.add_argument("--chroms", type=int, default=20)
    ap.add_argument("--repeat", type=int, default=5)
    args = ap.parse_args()
    import torch

    from gbrs_b200 import _lib
    from gbrs_b200 import reconstruct as rc
    from gbrs_b200 import synth

    w = np.linspace(1.6, 0.5, args.chroms)
    per = np.maximum((w / w.sum() * args.genes).astype(int), 2)
    base = synth.generate_reconstruct(genes_per_chrom=tuple(int(x) for x in per), H=8, sample_index=0, empty_chrom=False)
    tables = [base.expr] + [synth.generate_reconstruct(genes_per_chrom=tuple(int(x) for x in per), H=8, sample_index=s,
                                                       empty_chrom=False).expr for s in range(1, min(args.samples, 4))]
    tables = [tables[s % len(tables)] for s in range(args.samples)]
    out = {"metric": "reconstruct_gene_steps_per_s", "unit": "gene-steps/s (one gene of one sample: forward + backward "
           "+ Viterbi over 36 x 36 transitions)", "genes": int(per.sum()), "chromosomes": args.chroms, "H": 8, "S": 36}
    lib = _lib.load()
    for label, n_s in (("one_sample", 1), ("cohort", args.samples)):
        plan = rc.build_plan(base.chroms, base.genes, base.tprob, base.avecs, tables[:n_s], 8)
        t0 = time.perf_counter()
        res = rc.run_plan_on_device(plan, 1.5, 0.12)
        torch.cuda.synchronize()
        first = time.perf_counter() - t0
        e2e = []
        for _ in range(args.repeat):
            t0 = time.perf_counter()
            res = rc.run_plan_on_device(plan, 1.5, 0.12)
            torch.cuda.synchronize()
            e2e.append(time.perf_counter() - t0)
        # device-timed: resident tables, events around the two launches
        dev = torch.device("cuda", torch.cuda.current_device())
        S, G = plan.S, plan.expr.shape[0]

        def up(a):
            return torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1)).to(dev)

        d = {k: up(getattr(plan, k)) for k in ("expr", "avec_index", "avecs", "init", "tprob")}
        d["chains"] = up(plan.launch_order())
        buf = {k: torch.empty((G, S), dtype=torch.float64, device=dev) for k in ("eprob", "alpha", "gamma", "delta")}
        scaler = torch.empty(G, dtype=torch.float64, device=dev)
        n_mat = int(plan.tprob.shape[0])
        tlin = torch.empty(max(n_mat * S * S, 1), dtype=torch.float64, device=dev)
        backptr = torch.zeros((G, S), dtype=torch.uint8, device=dev)
        states = torch.zeros(plan.n_states_out, dtype=torch.int32, device=dev)
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        ms_e, ms_c = [], []
        for _ in range(args.repeat + 1):
            ev[0].record()
            _lib.check(lib.gbrs_hmm_emission(G, 8, d["expr"].data_ptr(), d["avecs"].data_ptr(), d["avec_index"].data_ptr(),
                                             d["init"].data_ptr(), 1.5, 0.12, buf["eprob"].data_ptr(), stream))
            ev[1].record()
            _lib.check(lib.gbrs_hmm_run(len(plan.chains), d["chains"].data_ptr(), 8, d["init"].data_ptr(),
                                        buf["eprob"].data_ptr(), d["tprob"].data_ptr(), n_mat, tlin.data_ptr(),
                                        buf["alpha"].data_ptr(),
                                        scaler.data_ptr(), buf["gamma"].data_ptr(), buf["delta"].data_ptr(),
                                        backptr.data_ptr(), states.data_ptr(), stream))
            ev[2].record()
            torch.cuda.synchronize()
            ms_e.append(ev[0].elapsed_time(ev[1]))
            ms_c.append(ev[1].elapsed_time(ev[2]))
        ms_e, ms_c = float(np.median(ms_e[1:])), float(np.median(ms_c[1:]))
        assert np.array_equal(states.cpu().numpy(), res["states"])
        steps = G
        mat_bytes = 8 * S * S * int(sum(len(base.tprob[c]) for c in plan.chroms))
        table_bytes = G * S * (8 * 4 + 8 + 2) + 8 * G  # eprob r, alpha w+r, gamma w, delta w, backptr w+r, scaler
        algo = 3 * mat_bytes * n_s + 2 * mat_bytes + table_bytes  # forward, backward, Viterbi per sample + exp once
        out[label] = {"samples": n_s, "chains": int(len(plan.chains)), "emission_ms": ms_e, "chain_ms": ms_c,
                      "value": steps / ((ms_e + ms_c) * 1e-3), "e2e_seconds": float(np.median(e2e)),
                      "first_call_seconds": first, "e2e_value": steps / float(np.median(e2e)),
                      "chain_algorithmic_GB": algo / 1e9, "chain_GBps": algo / (ms_c * 1e-3) / 1e9,
                      "unique_matrix_GB": mat_bytes / 1e9}
    # The CPU figure beside it comes from `python -m tests.time_reconstruct_oracle` (the oracle is test infrastructure
    # and is only executed from tests/ and bench.py): 1 core, ~1.3e3 genes/s on the build container.
    print(json.dumps(out))


if __name__ == "__main__":
    main()
