#!/usr/bin/env python
"""bench.py -- headline benchmark of the multiway EM quantifier (BASELINE.json: EM iterations/s and alignment-nnz/s,
model 4, DO-mouse-scale shape, % of the HBM roofline).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload c2|c1|small] [--model 4]

A *step* is one EM update (E-step + M-step + convergence test) over the resident packed incidence matrix.
  value     nnz * K / t          device-timed (CUDA events on the launch stream), inputs resident in HBM
  e2e       same metric through the public `EMfactory` API from HOST buffers: `EMfactory(apm).prepare()` [H2D of the H CSC
            matrices from pinned memory, packing on the device, theta0] + run(K iterations) + D2H of the expected counts, all
            inside the timed region; `e2e_resident` is the same with the already packed arrays (what round 1 called e2e)
  roofline  dominant kernel (the longer of the row and the column pass): algorithmic bytes of that kernel / its mean device
            time, vs MEASURED_PEAKS.json; `roofline.iteration`: the whole update by SURVEY 8(d)'s single-copy formula
  cpu_baseline  the UNMODIFIED reference (its own AlignmentPropertyMatrix + EMfactory, staged byte for byte into the
            git-ignored oracle/_ref/ by oracle/make_ref.py, `kind: "reference"`) timed on a bounded sample of the same
            workload on one host core; the oracle port (`kind: "port"`) only where the reference cannot run (models 1-3 on
            current scipy, or no staged copy)
  models    (default line only) models 3, 2, 1 on the same resident pattern, 5 updates each (BASELINE config 3)

N > 1 (torchrun): weak scaling -- every rank holds its own `workload`-sized row shard (different classes, same loci); the
T x 8 numerator is summed over the ranks once per step inside our own kernels over NVLink peer memory (GBRS_XCHG=push, the
default, | tag | pull | nvls; nccl: a plain all-reduce); time = max over ranks.  Every line carries a `parity` record (conservation, theta identical on all
ranks, a small problem sharded over the same ranks against the oracle).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    "c2": dict(T=80000, N=5_000_000, label="DO-mouse-scale quantify -M 4: 80k transcripts x 8 haplotypes x 5M classes"),
    "c1": dict(T=2000, N=200_000, label="config 1: 2k transcripts x 8 haplotypes x 200k classes"),
    "small": dict(T=8000, N=500_000, label="reduced: 8k x 8 x 500k (debug only)"),
    # BASELINE config 4: diploid (-G) restriction, 20 M classes in total, row-sharded (STRONG scaling: N / world each)
    "c4": dict(T=80000, N=20_000_000, diploid=True, strong=True,
               label="diploid quantify -G -M 4: 80k transcripts x 2 of 8 haplotypes per locus x 20M classes, row-sharded"),
    "c4small": dict(T=8000, N=2_000_000, diploid=True, strong=True, label="reduced diploid (debug only)"),
    # BASELINE config 5: batched cohort, 96 independent samples' EMs dealt to the GPUs (no communication)
    "c5": dict(T=80000, N=1_000_000, cohort=96, strong=True,
               label="batched cohort: 96 independent DO samples (80k transcripts x 8 haplotypes x 1M classes each), multiway "
                     "EM to convergence (tol 1e-4), samples dealt round-robin to the GPUs"),
    "c5small": dict(T=4000, N=100_000, cohort=16, strong=True, label="reduced cohort (debug only)"),
}
CPU_SAMPLE_CLASSES = 1_000_000
METRIC = "em_alignment_nnz_per_s"
UNIT = "nnz/s"


_REAL_STDOUT = None


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def emit(line: dict) -> None:
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, data)
    else:
        sys.stdout.write(data.decode())
        sys.stdout.flush()


# --------------------------------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.path = tempfile.mktemp(prefix="clocks_", suffix=".csv")
        self.proc = None
        try:
            self.fh = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={gpu_index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.fh,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.fh.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                it = [x.strip() for x in line.split(",")]
                if len(it) < 9:
                    continue
                try:
                    sm.append(float(it[1]))
                    mx.append(float(it[2]))
                except ValueError:
                    continue
                for nm, v in zip(names, it[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), samples=len(sm))
        out["reasons"] = sorted(reasons)
        return out


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port on a bounded sample
# --------------------------------------------------------------------------------------------------------------------
def cpu_port_rate(T, n_classes, steps, warmup, model=4, diploid=False):
    """nnz/s of the numpy restatement of the reference EM (oracle/em_oracle.py), single host thread."""
    from gbrs_b200 import synth
    from oracle import em_oracle as eo

    d = synth.generate(T=T, N=n_classes, H=8, with_genotype=diploid)
    apm = eo.apm_from_pairs(d.T, d.H, d.N, d.pair_class, d.pair_locus, d.pair_mask, d.count)
    if diploid:
        apm = eo.apply_genotype_mask(apm, synth.genotype_mask(d))
    eff = eo.effective_length_table(d.lengths)
    gene_of = eo.gene_index(d.T, d.groups()) if model != 4 else None
    keys = eo._Keys(apm, gene_of) if model != 4 else None
    theta = eo.prepare(apm, eff, 0.0)

    def one(theta):
        prev = theta.sum(axis=0)
        prev *= 1e6 / prev.sum()
        val = eo.e_step(apm, theta, model, gene_of, keys)
        theta = eo.sum_read(apm, val) / eff
        cur = theta.sum(axis=0)
        cur *= 1e6 / cur.sum()
        return theta, float(np.abs(cur - prev).sum())

    for _ in range(warmup):
        theta, _ = one(theta)
    t0 = time.perf_counter()
    for _ in range(steps):
        theta, _ = one(theta)
    dt = time.perf_counter() - t0
    return apm.nnz * steps / dt, dt / steps, apm.nnz


def cpu_reference_rate(T, n_classes, steps, warmup, diploid=False):
    """nnz/s of the UNMODIFIED reference (model 4): its own AlignmentPropertyMatrix + EMfactory, imported from the copy
    oracle/make_ref.py staged into oracle/_ref/ (or from /root/reference where that is mounted), driven through its
    public API: prepare(), then update_allelic_expression(model=4) per step -- the body of EMfactory.run's loop
    (EMfactory.py:267-279).  Single host thread (scipy / numpy sparse loops).  None if the reference is not available."""
    from oracle import ref_harness as rh

    if not rh.reference_available():
        return None
    from gbrs_b200 import synth

    d = synth.generate(T=T, N=n_classes, H=8, with_genotype=diploid)
    rh.install_shim(False)
    apm = rh.build_reference_apm(d, masked=diploid)
    nnz = int(sum(m.nnz for m in apm.data))
    with tempfile.TemporaryDirectory() as tmp:
        lenfile = os.path.join(tmp, "len.tsv")
        synth.write_length_file(d, lenfile)
        em = rh.load_reference().EMfactory(apm)
        t0 = time.perf_counter()
        em.prepare(pseudocount=0.0, lenfile=lenfile)
        t_prepare = time.perf_counter() - t0
    for _ in range(warmup):
        em.update_allelic_expression(model=4)
    t0 = time.perf_counter()
    for _ in range(steps):
        em.update_allelic_expression(model=4)
    dt = time.perf_counter() - t0
    return nnz * steps / dt, dt / steps, nnz, t_prepare


def cpu_baseline(T, n_classes, steps, warmup, model, diploid):
    """(rate, s_per_step, nnz, kind, note): the reference itself for model 4, the oracle port otherwise."""
    if model == 4:
        try:
            r = cpu_reference_rate(T, n_classes, steps, warmup, diploid)
        except Exception as e:  # noqa: BLE001 - e.g. a missing dependency of the reference on this host
            log(f"reference arm: the staged reference could not run ({type(e).__name__}: {e}); using the oracle port")
            r = None
        if r is not None:
            return r[0], r[1], r[2], "reference", f"unmodified reference EMfactory.update_allelic_expression(model=4); prepare() took {r[3]:.1f} s"
    rate, s_per, nnz = cpu_port_rate(T, n_classes, steps, warmup, model, diploid)
    why = "models 1-3 of the reference crash on current scipy (SURVEY.md fact 3)" if model != 4 else "oracle/_ref not staged"
    return rate, s_per, nnz, "port", f"oracle/em_oracle.py (numpy restatement of the reference; {why})"


def run_reference_arm(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # the whole workload (same config as our arm), a handful of updates: ~3 s per update for the reference at C2
    ncls = wl["N"] if args.model == 4 and wl["N"] <= 5_000_000 else min(CPU_SAMPLE_CLASSES, wl["N"])
    steps, warm = max(1, min(args.steps, 3)), min(args.warmup, 1)
    rate, s_per_step, nnz, kind, note = cpu_baseline(wl["T"], ncls, steps, warm, args.model, bool(wl.get("diploid")))
    sample = f"{ncls} of {wl['N']} classes (nnz={nnz}), {steps} EM updates timed after {warm} warm-up; {note}"
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warm, "ms_per_step": s_per_step * 1e3, "higher_is_better": True,
            "scaling": "strong" if wl.get("strong") else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl["label"], "model": args.model, "sample": sample},
            "iterations_per_s_at_full_size": rate / (nnz * wl["N"] / ncls),
            "cpu_baseline": {"value": rate, "unit": UNIT, "cores": 1, "kind": kind, "sample": sample},
            "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# --------------------------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------------------------
def pin_apm(apm):
    """Move the index arrays and counts of the host CSC matrices into pinned memory (in place)."""
    import torch

    for m in apm.data:
        for name in ("indices", "indptr"):
            a = getattr(m, name)
            setattr(m, name, torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy())
    if apm.count is not None:
        apm.count = torch.from_numpy(np.ascontiguousarray(apm.count, dtype=np.float64)).pin_memory().numpy()


def parity_checks(em, pat, d, world, rank, dev, diploid, model):
    """Checks of the path that was just timed, at every world size (the JSON line carries the result):
      conservation   sum of the expected counts == sum of the class counts over all ranks (every class' posterior sums to one)
      theta          bit-identical on all ranks (element-wise MIN == MAX over ranks)
      oracle         a small seeded problem, row-sharded over the same ranks through the same exchange, against the
                     oracle on the union of the shards (6 fixed updates, model 4 and the timed model)"""
    import torch
    import torch.distributed as dist

    from gbrs_b200 import synth
    from gbrs_b200.emfactory import EMfactory
    from oracle import em_oracle as eo

    out = {}
    counts = torch.from_numpy(np.ascontiguousarray(em.expected_read_counts())).to(dev)
    tot = torch.tensor([float(counts.sum().item()), float(d.count.sum()) if not diploid else 0.0], dtype=torch.float64, device=dev)
    theta = torch.from_numpy(np.ascontiguousarray(em.get_allelic_expression())).to(dev)
    lo, hi = theta.clone(), theta.clone()
    if world > 1:
        # every rank holds the summed numerator already: the count total is per-rank data, the expected total is not
        t2 = tot[1:].clone()
        dist.all_reduce(t2)
        tot[1] = t2[0]
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    out["theta_identical_across_ranks"] = bool(torch.equal(lo, hi))
    if not diploid:  # (with the -G restriction classes can lose all their alignments: the totals differ by design)
        out["conservation_rel"] = abs(float(tot[0].item()) - float(tot[1].item())) / float(tot[1].item())
    # small seeded problem through the same sharding
    Ts, Ns = 2000, 40_000
    ds = synth.generate(T=Ts, N=Ns, H=8, sample_index=100 + rank)
    ems = EMfactory(synth.to_apm(ds), device=dev, shard="local" if world > 1 else None)
    ems.target_lengths = synth.effective_lengths(ds)
    ems.prepare()
    rel = 0.0
    if rank == 0:
        parts = [synth.generate(T=Ts, N=Ns, H=8, sample_index=100 + r) for r in range(world)]
        pc = np.concatenate([p.pair_class + r * Ns for r, p in enumerate(parts)])
        oapm = eo.apm_from_pairs(Ts, 8, Ns * world, pc, np.concatenate([p.pair_locus for p in parts]),
                                 np.concatenate([p.pair_mask for p in parts]), np.concatenate([p.count for p in parts]))
        eff = eo.effective_length_table(parts[0].lengths)
        gene_of = eo.gene_index(Ts, parts[0].groups())
        want = eo.prepare(oapm, eff, 0.0)
        rel = max(rel, float(np.abs(ems.get_allelic_expression() - want).max() / np.abs(want).max()))
    for m in sorted({4, model}):
        ems.run(model=m, tol=0.0, max_iters=6, verbose=False)
        if rank == 0:
            o = eo.run(oapm, want, m, eff, gene_of, tol=0.0, max_iters=6)
            want = o["theta"]
            rel = max(rel, float(np.abs(ems.get_allelic_expression() - want).max() / np.abs(want).max()))
            rel = max(rel, float(np.abs(ems.expected_read_counts() - o["counts"]).max() / np.abs(o["counts"]).max()))
    out["oracle_small_shard_relerr"] = rel
    ok = out["theta_identical_across_ranks"] and out.get("conservation_rel", 0.0) < 1e-9 and rel < 1e-9
    if world > 1:
        flag = torch.tensor([1 if (ok or rank != 0) else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        ok = bool(flag.item())
    out["ok"] = bool(ok)
    assert ok, f"parity check failed: {out}"
    return out


def run_cohort_arm(args, wl):
    """BASELINE config 5.  A step = one whole sample: H2D of its CSC matrices (pinned host memory), packing on the device,
    prepare, EM to convergence, D2H of theta and the expected counts -- through gbrs_b200.cohort.quantify_cohort.  The
    samples are resident in host memory when the timed region starts (reading alignment files is not part of the metric);
    each rank generates a few distinct samples and cycles through them."""
    import torch
    import torch.distributed as dist

    from gbrs_b200 import cohort, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    n_samples = wl["cohort"]
    mine = cohort.my_share(n_samples, rank, world)
    distinct = min(4, len(mine))
    apms = []
    base = synth.generate(T=wl["T"], N=1000, H=8)
    eff = synth.effective_lengths(base)  # shared by the cohort (same transcriptome)
    for j in range(distinct):
        dj = synth.generate(T=wl["T"], N=wl["N"], H=8, sample_index=rank * distinct + j)
        a = synth.to_apm(dj)
        pin_apm(a)
        apms.append(a)
    slot = {i: k % distinct for k, i in enumerate(mine)}

    def load(i):
        return apms[slot[i]]

    def run_once(stats):
        return cohort.quantify_cohort(list(range(n_samples)), load, model=args.model, target_lengths=eff, rank=rank,
                                      world=world, device=dev, stats=stats)

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    run_once({})  # warm-up pass over the whole share (graph instantiation, allocator pools, pinned staging)
    sync_all()
    stats = {}
    t0 = time.perf_counter()
    res = run_once(stats)
    torch.cuda.synchronize(dev)
    wall = time.perf_counter() - t0
    tt = torch.tensor([wall, stats["gpu_phase_s"], float(stats["nnz_iters"]), float(sum(r["iters"] for r in res.values()))],
                      dtype=torch.float64, device=dev)
    if world > 1:
        mx = tt.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(tt, op=dist.ReduceOp.SUM)
        wall = float(mx[0].item())
    busy = float(tt[1].item()) / (world * wall)
    clocks = sampler.stop() if sampler else None
    # parity: the first sample of rank 0 against the oracle (same iteration count, counts to 1e-9)
    parity = None
    if rank == 0:
        from oracle import em_oracle as eo

        d0 = synth.generate(T=wl["T"], N=wl["N"], H=8, sample_index=0)
        oapm = eo.apm_from_pairs(d0.T, d0.H, d0.N, d0.pair_class, d0.pair_locus, d0.pair_mask, d0.count)
        o = eo.run(oapm, eo.prepare(oapm, eff, 0.0), args.model, eff,
                   eo.gene_index(d0.T, d0.groups()) if args.model != 4 else None, tol=1e-4)
        r0 = res[mine[0]]
        rel = float(np.abs(r0["counts"] - o["counts"]).max() / np.abs(o["counts"]).max())
        parity = {"iters": [int(r0["iters"]), int(o["iters"])], "counts_relerr": rel, "ok": r0["iters"] == o["iters"] and rel < 1e-9}
        assert parity["ok"], parity
        nnz_iters, iters = float(tt[2].item()), float(tt[3].item())
        line = {"metric": METRIC, "value": nnz_iters / wall, "unit": UNIT, "n_gpus": world, "steps": n_samples, "warmup": n_samples,
                "ms_per_step": wall / n_samples * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": wl["label"], "model": args.model, "samples": n_samples, "classes_per_sample": wl["N"],
                           "distinct_samples_per_rank": distinct, "l2_policy": "every sample's matrices stream from pinned host memory"},
                "samples_per_s": n_samples / wall, "em_updates_total": iters,
                "gpu_phase_fraction": busy,  # share of the wall time spent in device packing / run / fetch (host waits on the GPU)
                "e2e": {"value": nnz_iters / wall, "unit": UNIT, "h2d_bytes_per_step": float(sum(m.indices.nbytes + m.indptr.nbytes for m in apms[0].data) + apms[0].count.nbytes),
                        "d2h_bytes_per_step": 2.0 * 8 * 8 * wl["T"], "seconds": wall,
                        "what": "quantify_cohort over all samples: per sample H2D of the CSC matrices, device packing, prepare, "
                                "EM to convergence, D2H of theta and counts"},
                "parity": parity, "clocks": clocks, "gpu_launches": int(iters) * 4}
        emit(line)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def run_gpu_arm(args, wl):
    import torch
    import torch.distributed as dist

    from gbrs_b200 import _lib, synth
    from gbrs_b200.emfactory import EMfactory

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the EM has no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # NCCL prints its version banner to stdout otherwise
        dist.init_process_group("nccl", device_id=dev)
    if world != args.gpus:
        log(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}")

    sampler = ClockSampler(local_rank) if rank == 0 else None

    # ---- workload: every rank holds one `workload`-sized shard (weak scaling) ------------------------------------
    t0 = time.perf_counter()
    diploid, strong = bool(wl.get("diploid")), bool(wl.get("strong"))
    n_local = wl["N"] // world if strong else wl["N"]
    d = synth.generate(T=wl["T"], N=n_local, H=8, sample_index=rank, with_genotype=diploid)
    apm = synth.to_apm(d)
    hapmask = None
    if diploid:
        import importlib

        hapmask = importlib.import_module("gbrs_b200.quantify").hapmask_bytes(synth.genotype_mask(d))
    t_gen = time.perf_counter() - t0
    # the resident pattern of the device-timed run comes from the host packer (its arrays also feed `e2e_resident` and the
    # layout statistics); the `e2e` passes below build their own pattern the default way (device packer)
    em = EMfactory(apm, device=dev, shard="local" if world > 1 else None, locus_hapmask=hapmask, pack="host")
    em.target_lengths = synth.effective_lengths(d)  # same table prepare() would parse from a targets.info file
    t0 = time.perf_counter()
    em.prepare()  # gene tables + pack + upload + theta0
    t_first = time.perf_counter() - t0
    pat = em._pattern
    if args.model != 4:
        pat.ensure_full()  # the arrays only models 1-3 read are uploaded on demand
    lib, desc = pat.lib, pat.desc
    info = pat.info
    nnz_local = info["nnz"]
    nnz_total = nnz_local
    if world > 1:
        t = torch.tensor([nnz_local], dtype=torch.int64, device=dev)
        dist.all_reduce(t)
        nnz_total = int(t.item())
    log(f"[rank {rank}] generate {t_gen:.1f}s, pack {pat.packed.pack_seconds:.2f}s, first prepare {t_first:.2f}s, "
        f"classes={info['n_classes']} pairs={info['n_pairs']} nnz={nnz_local} items={info['n_items']}")

    model = args.model

    def exchange():
        em._exchange(pat)  # NCCL all-reduce, or nothing when the kernels exchange over NVLink peer memory themselves

    def step():
        _lib.check(lib.gbrs_em_launch_local(C.byref(desc), model, pat.stream()))
        exchange()
        _lib.check(lib.gbrs_em_launch_update(C.byref(desc), pat.stream()))

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    K, W = args.steps, max(args.warmup, 3)
    # Steps are queued from a CUDA graph holding CHUNK updates (kernels + the all-reduce), replayed on a side stream;
    # the remainder and the per-kernel profiling loop use plain launches.
    CHUNK = 10
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    use_graph = os.environ.get("GBRS_NO_GRAPH") is None
    with torch.cuda.stream(side):
        stream = pat.stream()
        # tol = 0 keeps the loop alive for exactly the number of updates we queue
        _lib.check(lib.gbrs_em_run_begin(C.byref(desc), 0.0, min(4 * (K + W) + 64, 60000), stream))
        graph = None
        if use_graph:
            if world > 1:
                scratch = torch.zeros(8, dtype=torch.float64, device=dev)
                dist.all_reduce(scratch)
            side.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=side):
                for _ in range(CHUNK):
                    step()
            # capture does not execute: nothing has run yet

        def run_steps(n):
            full = n // CHUNK if graph is not None else 0
            for _ in range(full):
                graph.replay()
            for _ in range(n - full * CHUNK):
                step()

        run_steps(W)
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run_steps(K)
        e1.record()
        side.synchronize()
        ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    sync_all()
    stream = pat.stream()
    ctrl, scal = pat.read_ctrl()
    assert ctrl[_lib.CTRL_ERROR] == 0, "non-finite value during the timed region"
    assert ctrl[_lib.CTRL_ITERS] == W + K, (ctrl[:6], W, K)
    value = nnz_total * K / (ms * 1e-3)

    # ---- per-kernel device time of the same K updates (events around the row / column pass) --------------------
    prof = C.c_void_p()
    _lib.check(lib.gbrs_prof_create(K, C.byref(prof)))
    for _ in range(K):
        _lib.check(lib.gbrs_em_launch_local_profiled(C.byref(desc), model, stream, prof))
        exchange()
        _lib.check(lib.gbrs_em_launch_update(C.byref(desc), stream))
    ms_row, ms_col, ms_acc, nrec = C.c_double(), C.c_double(), C.c_double(), C.c_int32()
    _lib.check(lib.gbrs_prof_read(prof, C.byref(ms_row), C.byref(ms_col), C.byref(ms_acc), C.byref(nrec)))
    lib.gbrs_prof_free(prof)
    row_ms, col_ms, acc_ms = ms_row.value / K, ms_col.value / K, ms_acc.value / K
    per_rank_ms = None
    if world > 1:  # the step lasts as long as the slowest rank's passes plus the exchange: show every rank's kernel times
        mine = torch.tensor([row_ms, col_ms, acc_ms], dtype=torch.float64, device=dev)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank_ms = [[round(float(x), 5) for x in t.tolist()] for t in allr]
    clocks = sampler.stop() if sampler else None

    # ---- algorithmic bytes -----------------------------------------------------------------
    Np, Pp, It, T = info["n_classes"], info["n_pairs"], info["n_items"], wl["T"]
    eb = info["entry_bytes"]
    widx = {4: Np, 3: info["n_runs"], 2: Pp, 1: 8 * info["n_runs"]}[model]
    Ee = info["n_entries"]
    bytes_row = 4 * Pp + 8 * Np + 8 * widx + 256 * T  # pair words, counts, weights out, subset tables in
    bytes_col = eb * Ee + 16 * It + 8 * widx + 64 * It  # entry words, item descriptors, weights in, item sums out
    bytes_iter = 4 * Pp + 4 * (Np + 1) + 8 * Np + 3 * 64 * T  # SURVEY 8(d), pair+mask layout, inputs once
    peak, peak_src = measured_peaks()
    tiled = pat.tiled is not None and model == 4
    if tiled:
        ti = pat.tiled.info
        # the fused tile kernel: every tile blob once, the subset-table rows of its loci in, one 64-byte partial per
        # (tile, locus) slot out
        bytes_tile = ti["blob_bytes"] + 16 * ti["n_tiles"] + (256 + 64) * ti["n_slots"]
        dom, dom_ms, dom_bytes = "k_tile_em (fused E-step + column reduce over tiles)", row_ms, bytes_tile
    elif row_ms >= col_ms:
        dom, dom_ms, dom_bytes = "k_weights_m%d (row pass)" % model, row_ms, bytes_row
    else:
        dom, dom_ms, dom_bytes = "k_column_reduce (column pass)", col_ms, bytes_col
    achieved = dom_bytes / (dom_ms * 1e-3) / 1e9
    traffic = None
    try:  # DRAM bytes of that kernel from the committed ncu --set full capture (same workload and model only)
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        if tj.get("workload") == args.workload and tj.get("model") == model and world == 1:
            traffic = tj.get(dom)
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": dom_bytes, "kernel_ms": dom_ms,
                "share_of_step": dom_ms / (ms / K),
                "per_kernel_ms": {("tile_pass" if tiled else "row_pass"): row_ms, "column_pass": col_ms, "locus_acc": acc_ms,
                                  "rest_of_step": max(ms / K - row_ms - col_ms - acc_ms, 0.0)},
                "per_rank_row_column_locus_ms": per_rank_ms,
                "iteration": {"algorithmic_bytes": bytes_iter, "achieved": bytes_iter / (ms / K * 1e-3) / 1e9,
                              "frac": bytes_iter / (ms / K * 1e-3) / 1e9 / peak,
                              "note": "SURVEY 8(d) pair+mask formula (each input once, theta in/out + lengths); the "
                                      "second (locus-major) copy and the weight vector are NOT counted"}}

    # ---- e2e through the public API, from HOST buffers ----------------------------------------------------------------
    # `e2e`: what a user's call sequence costs -- EMfactory(apm) on the host CSC matrices, prepare() (host packing of the
    # incidence, H2D of the packed arrays, theta0), run(K updates), expected_read_counts() (D2H).  `e2e_resident`: the
    # same without the host packer (the packed arrays are already in pinned host memory) -- what round 1 reported.
    for k in list(pat.host):
        pat.host[k] = pat.host[k].pin_memory()
    pin_apm(apm)  # the inputs of `e2e` are copied from pinned host memory
    Ke = K
    sync_all()

    def e2e_resident_once():
        t0 = time.perf_counter()
        pat.h2d_bytes = 0
        pat.upload()
        torch.cuda.synchronize(dev)
        t1 = time.perf_counter()
        em.reset()
        t2 = time.perf_counter()
        em.run(model=model, tol=0.0, max_iters=Ke, verbose=False)
        t3 = time.perf_counter()
        c = em.expected_read_counts()
        torch.cuda.synchronize(dev)
        t4 = time.perf_counter()
        return t4 - t0, c, {"h2d_s": t1 - t0, "prepare_s": t2 - t1, "run_s": t3 - t2, "fetch_s": t4 - t3}

    def e2e_full_once():
        t0 = time.perf_counter()
        em2 = EMfactory(apm, device=dev, shard="local" if world > 1 else None, locus_hapmask=hapmask)
        em2.target_lengths = em.target_lengths
        em2.prepare()  # pack on the host + H2D + theta0
        t1 = time.perf_counter()
        em2.run(model=model, tol=0.0, max_iters=Ke, verbose=False)
        t2 = time.perf_counter()
        c = em2.expected_read_counts()
        torch.cuda.synchronize(dev)
        t3 = time.perf_counter()
        p2 = em2._pattern
        parts = {"pack_s": p2.packed.pack_seconds, "packer": "device (gbrs_pack_device)" if p2.on_device else "host (gbrs_pack_create)",
                 "tiles_s": p2.tiled.build_seconds if p2.tiled is not None else 0.0,
                 "prepare_total_s": t1 - t0, "run_s": t2 - t1, "fetch_s": t3 - t2}
        return t3 - t0, c, parts, p2.h2d_bytes, em2.num_iters

    e2e_resident_once()  # warm-up pass (first-use costs: pinned staging, graph instantiation paths), then the timed one
    sync_all()
    t_res, counts, res_parts = e2e_resident_once()
    assert em.num_iters == Ke
    sync_all()
    e2e_full_once()
    sync_all()
    t_e2e, counts2, e2e_parts, h2d_full, iters2 = e2e_full_once()
    assert iters2 == Ke
    if world > 1:
        t = torch.tensor([t_e2e, t_res], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_e2e, t_res = float(t[0].item()), float(t[1].item())
    if world == 1 and not diploid:
        assert abs(counts.sum() - d.count.sum()) < 1e-6 * counts.sum()
        assert abs(counts2.sum() - d.count.sum()) < 1e-6 * counts2.sum()
    h2d_res = pat.h2d_bytes + 64 * T + 64 * T  # packed arrays + effective lengths (+ nothing else)
    d2h = 2 * 64 * T + 2 * 8 * wl["T"] * 8 + 8 * Ke
    e2e = {"value": nnz_total * Ke / t_e2e, "unit": UNIT, "h2d_bytes_per_step": (h2d_full + 128 * T) / Ke,
           "d2h_bytes_per_step": d2h / Ke, "seconds": t_e2e, "parts": e2e_parts,
           "what": "EMfactory(apm).prepare() [H2D of the CSC matrices from pinned memory, packing on the device, theta0] + "
                   "run(%d updates) + expected_read_counts() [D2H]" % Ke}
    e2e_resident = {"value": nnz_total * Ke / t_res, "unit": UNIT, "h2d_bytes_per_step": h2d_res / Ke,
                    "d2h_bytes_per_step": d2h / Ke, "seconds": t_res, "parts": res_parts,
                    "what": "H2D of the already packed incidence (pinned) + reset + run(%d updates) + D2H" % Ke}

    # ---- the other three models on the same resident pattern (BASELINE config 3) ------------------------------------------
    models_rec = None
    if model == 4 and world == 1 and args.workload == "c2" and not args.no_models:
        models_rec = {}
        pat.ensure_full()
        _lib.check(lib.gbrs_em_run_begin(C.byref(desc), 0.0, 1000, stream))  # re-arm the loop (the e2e run has stopped it)
        for m in (3, 2, 1):
            for _ in range(2):
                _lib.check(lib.gbrs_em_launch_local(C.byref(desc), m, stream))
                _lib.check(lib.gbrs_em_launch_update(C.byref(desc), stream))
            torch.cuda.synchronize(dev)
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            f0.record(torch.cuda.current_stream(dev))
            for _ in range(5):
                _lib.check(lib.gbrs_em_launch_local(C.byref(desc), m, stream))
                _lib.check(lib.gbrs_em_launch_update(C.byref(desc), stream))
            f1.record(torch.cuda.current_stream(dev))
            torch.cuda.synchronize(dev)
            mm = f0.elapsed_time(f1) / 5
            models_rec[str(m)] = {"ms_per_step": mm, "value": nnz_total / (mm * 1e-3), "unit": UNIT, "steps": 5}
        ctrl_m, _ = pat.read_ctrl()
        assert ctrl_m[_lib.CTRL_ERROR] == 0 and ctrl_m[_lib.CTRL_ITERS] == 3 * 7, ctrl_m[:6]

    # ---- parity of the timed path (every run carries it; at N > 1 this is where the sharded result gets checked) ------
    parity = parity_checks(em, pat, d, world, rank, dev, diploid, args.model)

    # ---- CPU baseline (rank 0, N=1 only): a bounded sample of the same workload on the host cores ---------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        ncls = min(CPU_SAMPLE_CLASSES, wl["N"])
        rate, s_per, nnz_s, kind, note = cpu_baseline(wl["T"], ncls, 5, 1, model, diploid)
        cpu = {"value": rate, "unit": UNIT, "cores": 1, "kind": kind,
               "sample": f"{ncls} of {wl['N']} classes (nnz={nnz_s}), 5 EM updates, {s_per:.3f} s/update; {note}; host has "
                         f"{os.cpu_count()} logical cores (the path is single-threaded)"}

    xphases = None
    if world > 1 and em.fused_exchange and em.exchange_mode in ("push", "tag"):
        st = pat.part[-8:-3].cpu().numpy() * 1e-3  # us since kernel start, block 0 of the last update
        if em.exchange_mode == "push":
            xphases = {"numerator_block0": float(st[0]), "all_ready": float(st[1]), "slice_reduced": float(st[2]),
                       "all_done": float(st[3]), "update_done": float(st[4])}
        else:  # no flags to wait for: block 0's first complete pair of its slice, its slice done, its share of the update done
            xphases = {"numerator_block0": float(st[0]), "first_pair_summed": float(st[1]), "slice_reduced": float(st[2]),
                       "update_done": float(st[4])}
            raw = pat.part[-8:].cpu().numpy()
            t0 = raw[3]
            late = raw[5:8].copy().view(np.uint64).astype(np.float64)  # latest block's end of phases A / B / C (absolute)
            mine = torch.tensor([(late[0] - t0) * 1e-3, (late[1] - t0) * 1e-3, (late[2] - t0) * 1e-3], dtype=torch.float64, device=dev)
            allr = [torch.zeros_like(mine) for _ in range(world)]
            dist.all_gather(allr, mine)
            xphases["slowest_block_A_B_C_per_rank"] = [[round(float(x), 2) for x in t.tolist()] for t in allr]
    if rank == 0:
        pushed = world > 1 and em.fused_exchange and em.exchange_mode in ("push", "tag")
        launches_per_step = (3 if tiled else 4) + (1 if world > 1 and not pushed else 0) + (1 if model != 4 else 0)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong" if strong else "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": wl["label"], "model": model, "classes_per_gpu": Np, "pairs_per_gpu": Pp,
                           "nnz_per_gpu": nnz_local, "nnz_total": nnz_total, "loci": T, "haplotypes": 8,
                           "layout": ("tiles: %d tiles, %.1f MB of blobs, %d partial slots" % (
                               pat.tiled.info["n_tiles"], pat.tiled.info["blob_bytes"] / 1e6, pat.tiled.info["n_slots"]))
                           if tiled else "two-pass: class-major + locus-major copies",
                           "l2_policy": "inputs larger than L2 (packed incidence %.0f MB per GPU streams every step)"
                                        % ((pat.tiled.nbytes() if tiled else pat.packed.nbytes()) / 1e6),
                           "exchange": "none" if world == 1 else {
                               "push": "one launch per update: local numerator pushed into the owners' receive rows over NVLink "
                                       "peer memory, slice sums broadcast by their owners, update (k_locus_xchg)",
                               "tag": "one launch per update, no flags: every double of the numerator carries the parity of its "
                                      "exchange in its sign bit and is polled by the thread that needs it (push to the owners' "
                                      "receive rows, slice sums broadcast, update -- k_locus_xchg<.., TAG>)",
                               "pull": "fused two-shot all-reduce of T x 8 fp64 over NVLink peer memory inside the locus kernels",
                               "nvls": "fused two-shot all-reduce of T x 8 fp64 inside the locus kernels: NVLS (multimem.ld_reduce / "
                                       "multimem.st on the NVSwitch multicast mapping)",
                               "nccl": "NCCL all-reduce of T x 8 fp64 per step"}[em.exchange_mode if em.fused_exchange else "nccl"]},
                "iterations_per_s": K / (ms * 1e-3), "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
                "e2e_resident": e2e_resident, "parity": parity, "models": models_rec, "exchange_phases_us": xphases,
                "gpu_launches": launches_per_step * K, "clocks": clocks,
                "pack_seconds": e2e_parts["pack_s"], "pack_seconds_host_packer": pat.packed.pack_seconds}
        emit(line)
    # teardown: drop the captured graph (it holds NCCL work) before the communicator, and never hang on exit
    graph = None
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def main():
    # Keep stdout to the one JSON line: libraries (NCCL's version banner) write to fd 1, so fd 1 is pointed at stderr
    # for the whole run and the line is written to the original stdout at the end.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--model", type=int, default=4, choices=[1, 2, 3, 4])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-models", action="store_true", help="skip the models 1-3 sub-record of the default line")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference_arm(args, wl)
    elif wl.get("cohort"):
        run_cohort_arm(args, wl)
    else:
        run_gpu_arm(args, wl)


if __name__ == "__main__":
    main()
