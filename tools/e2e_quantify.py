#!/usr/bin/env python
"""End-to-end `quantify` at benchmark scale: writes a synthetic EMASE file + group / length (/ genotype) files, runs
`gbrs_b200.quantify.quantify` exactly as the CLI would, and prints the wall time of every phase as one JSON line.

    python tools/e2e_quantify.py [--T 80000] [--N 5000000] [--model 4] [--diploid] [--keep DIR]
"""
import argparse
import importlib
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--T", type=int, default=80000)
    ap.add_argument("--N", type=int, default=5_000_000)
    ap.add_argument("--model", type=int, default=4)
    ap.add_argument("--diploid", action="store_true")
    ap.add_argument("--keep", default=None)
    ap.add_argument("--profile", action="store_true", help="cProfile of the quantify() call on stderr")
    args = ap.parse_args()

    import numpy as np

    from gbrs_b200 import emfactory, synth

    qmod = importlib.import_module("gbrs_b200.quantify")
    times = {}
    t0 = time.perf_counter()
    d = synth.generate(T=args.T, N=args.N, H=8, with_genotype=True)
    apm = synth.to_apm(d)
    times["generate_s"] = time.perf_counter() - t0
    tmp = args.keep or tempfile.mkdtemp(prefix="gbrs_e2e_")
    os.makedirs(tmp, exist_ok=True)
    aln, grp, ln, gt = (os.path.join(tmp, x) for x in ("aln.emase", "grp.tsv", "len.tsv", "gt.tsv"))
    t0 = time.perf_counter()
    apm.save(aln)
    synth.write_group_file(d, grp)
    synth.write_length_file(d, ln)
    synth.write_genotype_file(d, gt)
    times["write_inputs_s"] = time.perf_counter() - t0

    # phase timing through light instrumentation of the public classes
    marks = {}

    def timed(obj, name, key):
        orig = getattr(obj, name)

        def wrapper(*a, **k):
            t = time.perf_counter()
            try:
                return orig(*a, **k)
            finally:
                marks[key] = marks.get(key, 0.0) + time.perf_counter() - t
        setattr(obj, name, wrapper)

    timed(qmod, "AlignmentPropertyMatrix", "load_file_s")
    timed(emfactory.EMfactory, "prepare", "prepare_total_s")
    timed(emfactory.EMfactory, "run", "run_s")
    timed(emfactory.EMfactory, "report_depths", "reports_s")
    timed(emfactory.EMfactory, "report_read_counts", "reports_s")
    timed(emfactory.PackedPattern, "__init__", "pack_host_s")
    timed(emfactory.DevicePacked, "__init__", "pack_device_s")
    timed(emfactory.EMfactory, "_read_lengths", "read_lengths_s")
    import torch

    t0 = time.perf_counter()
    torch.cuda.init()
    torch.zeros(1, device="cuda")
    torch.cuda.synchronize()
    times["cuda_context_s"] = time.perf_counter() - t0  # paid once per process, before quantify() is entered
    import cProfile
    import io
    import pstats

    pr = cProfile.Profile() if args.profile else None
    t0 = time.perf_counter()
    if pr:
        pr.enable()
    qmod.quantify(alignment_file=aln, group_file=grp, length_file=ln, genotype_file=gt if args.diploid else None,
                  outbase=os.path.join(tmp, "out"), multiread_model=args.model, report_alignment_counts=True)
    if pr:
        pr.disable()
        buf = io.StringIO()
        pstats.Stats(pr, stream=buf).sort_stats("cumulative").print_stats(35)
        print(buf.getvalue()[:7000], file=sys.stderr)
    times["quantify_total_s"] = time.perf_counter() - t0
    times.update(marks)
    kind = "diploid" if args.diploid else "multiway"
    sizes = {f: os.path.getsize(os.path.join(tmp, f)) for f in sorted(os.listdir(tmp)) if f.startswith("out.")}
    tab = np.loadtxt(os.path.join(tmp, f"out.{kind}.isoforms.expected_read_counts"), skiprows=1, usecols=range(1, 10))
    times["sum_expected_counts"] = float(tab[:, -1].sum())
    times["sum_class_counts"] = float(d.count.sum())
    print(json.dumps({"T": args.T, "N": args.N, "model": args.model, "diploid": args.diploid, "times": times,
                      "outputs": sizes}))


if __name__ == "__main__":
    main()
