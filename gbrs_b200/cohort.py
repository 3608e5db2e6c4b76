"""Batched cohort mode (BASELINE.json config 5): many independent per-sample multiway EMs.

Replicas only: samples do not interact, so they are dealt round-robin to the ranks of a
torch.distributed job (or all run on the one local GPU) and each is quantified by its own `EMfactory`; there is no
collective on the data path.  What a sample costs the GPU is milliseconds (packing on the device + a converged EM);
what it costs the host -- reading and inflating the alignment file -- is far more, so the samples of a rank form a
pipeline: a loader thread fetches sample i+1 (the user's `load` callback: file read / inflate / pinning release the
GIL) while the device packs and runs sample i.  The effective-length table and the gene grouping are shared by the
samples of a cohort (same transcriptome): they are parsed once.  Device buffers of a finished sample are released
before the next one is packed, so any number of samples fits.
"""
from __future__ import annotations

import queue
import threading
import time
from typing import Callable, Iterable, Sequence

from . import utils
from .emfactory import EMfactory

logger = utils.get_logger("gbrs")


def my_share(n_samples: int, rank: int = 0, world: int = 1) -> list[int]:
    """Indices of the samples rank `rank` of `world` works on (round-robin)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank / world")
    return list(range(rank, n_samples, world))


def _prefetch(indices, samples, load, depth):
    """Yield (index, apm, seconds spent loading) with up to `depth` samples loaded ahead by a background thread."""
    if depth <= 0:
        for i in indices:
            t0 = time.perf_counter()
            apm = load(samples[i])
            yield i, apm, time.perf_counter() - t0
        return
    q: queue.Queue = queue.Queue(maxsize=depth)

    def worker():
        try:
            for i in indices:
                t0 = time.perf_counter()
                apm = load(samples[i])
                apm.is_pure_incidence(cache=True)  # a pass over the stored values: done here, off the device's critical path
                q.put((i, apm, time.perf_counter() - t0))
        except BaseException as e:  # noqa: BLE001 - re-raised by the consumer
            q.put(e)
        q.put(None)

    th = threading.Thread(target=worker, name="gbrs-cohort-loader", daemon=True)
    th.start()
    while True:
        item = q.get()
        if item is None:
            break
        if isinstance(item, BaseException):
            raise item
        yield item
    th.join()


def quantify_cohort(samples: Sequence, load: Callable, model: int = 4, pseudocount: float = 0.0,
                    lenfile: str | None = None, read_length: int = 100, tol: float = 0.0001, max_iters: int = 999,
                    device=None, rank: int | None = None, world: int | None = None,
                    on_done: Callable | None = None, prefetch: int = 2, stats: dict | None = None,
                    target_lengths=None) -> dict:
    """Run the EM of every sample this rank owns.

    samples   any sequence of sample descriptors (file names, ids ...)
    load      callable(sample) -> AlignmentPropertyMatrix (groups attached if gene-level output is wanted)
    on_done   optional callable(sample, EMfactory) invoked after each sample (e.g. to write its report files)
    prefetch  samples loaded ahead by the loader thread (0: load in line)
    stats     optional dict, filled with the wall seconds of this rank's share, the seconds spent per sample in total
              (`device_s`), in the phases where the host only waits for the GPU (`gpu_phase_s`: packing, run, fetch) and
              in `load`
    target_lengths  optional H x T effective-length table (instead of `lenfile`), shared by all samples
    Returns {sample index: dict(theta=H x T depths, counts=H x T expected read counts, iters=int)}.
    """
    if rank is None or world is None:
        rank, world = 0, 1
        try:
            import torch.distributed as dist

            if dist.is_available() and dist.is_initialized():
                rank, world = dist.get_rank(), dist.get_world_size()
        except Exception:  # noqa: BLE001
            pass
    import torch

    out = {}
    shared_lengths = target_lengths  # H x T effective lengths; else parsed from `lenfile` by the first sample
    t_wall = time.perf_counter()
    t_load = t_dev = t_gpu = 0.0
    nnz_iters = 0
    shared_groups = None
    for i, apm, dt_load in _prefetch(my_share(len(samples), rank, world), samples, load, prefetch):
        t_load += dt_load
        t0 = time.perf_counter()
        # one grouping for the cohort: equal lists become ONE list object, so that the gene tables derived from it are
        # built once (utils._cached)
        if shared_groups is not None and apm.groups is not shared_groups and apm.groups == shared_groups:
            apm.groups = shared_groups
        elif shared_groups is None and getattr(apm, "groups", None):
            shared_groups = apm.groups
        em = EMfactory(apm, device=device)  # no process group: the sample is not sharded
        if shared_lengths is not None:
            em.target_lengths = shared_lengths
            em.prepare(pseudocount=pseudocount)
        else:
            em.prepare(pseudocount=pseudocount, lenfile=lenfile, read_length=read_length)
            shared_lengths = em.target_lengths
        t1 = time.perf_counter()
        em.run(model=model, tol=tol, max_iters=max_iters, verbose=False)
        out[i] = dict(theta=em.get_allelic_expression(), counts=em.expected_read_counts().copy(), iters=em.num_iters)
        torch.cuda.synchronize(em._pattern.device)
        t_dev += time.perf_counter() - t0
        # phases in which the host only waits for the device: packing (H2D + kernels, ends with a sync) and run + fetch
        t_gpu += getattr(getattr(em._pattern, "packed", None), "pack_seconds", 0.0) + (time.perf_counter() - t1)
        nnz_iters += em._pattern.info["nnz"] * em.num_iters
        if on_done is not None:
            on_done(samples[i], em)
        logger.info(f"sample {i}: {em.num_iters} EM updates")
        del em
    if stats is not None:
        stats.update(wall_s=time.perf_counter() - t_wall, device_s=t_dev, gpu_phase_s=t_gpu, load_s=t_load, samples=len(out),
                     nnz_iters=nnz_iters)
    return out


def gather_results(local: dict, n_samples: int) -> Iterable:
    """Collect the per-rank dictionaries on every rank (object all-gather); single process: identity."""
    try:
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            parts = [None] * dist.get_world_size()
            dist.all_gather_object(parts, local)
            merged = {}
            for p in parts:
                merged.update(p)
            return [merged[i] for i in range(n_samples)]
    except ImportError:
        pass
    return [local[i] for i in range(n_samples)]
