"""Batched cohort mode (BASELINE.json config 5): many independent per-sample multiway EMs.

Replicas only: samples do not interact, so they are dealt round-robin to the ranks of a
torch.distributed job (or all run on the one local GPU) and each is quantified by its own `EMfactory`; there is no
collective on the data path.  Device buffers of a finished sample are released before the next one is packed, so any
number of samples fits.
"""
from __future__ import annotations

from typing import Callable, Iterable, Sequence

from . import utils
from .emfactory import EMfactory

logger = utils.get_logger("gbrs")


def my_share(n_samples: int, rank: int = 0, world: int = 1) -> list[int]:
    """Indices of the samples rank `rank` of `world` works on (round-robin)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank / world")
    return list(range(rank, n_samples, world))


def quantify_cohort(samples: Sequence, load: Callable, model: int = 4, pseudocount: float = 0.0,
                    lenfile: str | None = None, read_length: int = 100, tol: float = 0.0001, max_iters: int = 999,
                    device=None, rank: int | None = None, world: int | None = None,
                    on_done: Callable | None = None) -> dict:
    """Run the EM of every sample this rank owns.

    samples   any sequence of sample descriptors (file names, ids ...)
    load      callable(sample) -> AlignmentPropertyMatrix (groups attached if gene-level output is wanted)
    on_done   optional callable(sample, EMfactory) invoked after each sample (e.g. to write its report files)
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
    out = {}
    for i in my_share(len(samples), rank, world):
        apm = load(samples[i])
        em = EMfactory(apm, device=device)  # no process group: the sample is not sharded
        em.prepare(pseudocount=pseudocount, lenfile=lenfile, read_length=read_length)
        em.run(model=model, tol=tol, max_iters=max_iters, verbose=False)
        out[i] = dict(theta=em.get_allelic_expression(), counts=em.expected_read_counts().copy(), iters=em.num_iters)
        if on_done is not None:
            on_done(samples[i], em)
        logger.info(f"sample {i}: {em.num_iters} EM updates")
        del em
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
