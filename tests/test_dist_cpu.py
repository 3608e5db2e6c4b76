"""world_size-2 gloo test of the row-sharded N > 1 host logic on CPU: each rank packs only its own shard (C++ packer),
computes its local numerator by walking the packed arrays (tests/packed_emulation.py -- the numpy mirror of the
kernels), the T x 8 numerator is summed with torch.distributed (gloo) exactly where EMfactory._exchange sits, and every
rank finishes the update.  All ranks must agree bit for bit and match the reference golden trajectory."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
sys.path.insert(0, os.environ["GBRS_ROOT"])
import numpy as np, torch, torch.distributed as dist
from gbrs_b200 import synth
from gbrs_b200.emfactory import PackedPattern, EMfactory
from oracle import em_oracle as eo
from tests import helpers as hp
from tests.packed_emulation import em_update_model4

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
g = hp.load_golden("em_small_m4")
d = hp.synth_from_golden(g)
apm = synth.to_apm(d)
# EMfactory resolves rank / world from the process group without touching the GPU
em = EMfactory(apm, shard=True)
assert (em.rank, em.world) == (rank, world)
p = PackedPattern(apm, shard_rank=rank, shard_count=world)
nnz = torch.tensor([p.info["nnz"]]); dist.all_reduce(nnz)
assert int(nnz) == d.nnz == p.info["nnz_total"]
eff = np.ones((d.T, 8)); eff[:, :d.H] = eo.effective_length_table(d.lengths).T

def update(theta, unit=False):
    acc, _ = em_update_model4(p.arrays, p.info, d.T, theta, eff, unit=unit)
    t = torch.from_numpy(acc); dist.all_reduce(t)          # the one exchange step per update
    return t.numpy(), t.numpy() / eff

acc, theta = update(None, unit=True)
assert hp.relerr(theta[:, :d.H].T, g["theta0"]) < 1e-12
errs = []
for _ in range(g["iters"]):
    prev = theta.sum(axis=1); prev = prev * (1e6 / prev.sum())
    acc, theta = update(theta)
    cur = theta.sum(axis=1); cur = cur * (1e6 / cur.sum())
    errs.append(np.abs(cur - prev).sum())
assert errs[-1] <= 1e6 * g["tol"] < errs[-2]               # same stop decision as the reference
assert hp.relerr(theta[:, :d.H].T, g["theta"]) < 1e-10 and hp.relerr(acc[:, :d.H].T, g["counts"]) < 1e-10
gathered = [torch.zeros(d.T, 8, dtype=torch.float64) for _ in range(world)]
dist.all_gather(gathered, torch.from_numpy(theta))
assert all(torch.equal(x, gathered[0]) for x in gathered)  # identical on every rank -> identical decisions
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_two_rank_gloo_sharded_update(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, GBRS_ROOT=ROOT, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="2")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", "29513", str(script)]
    res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert res.stdout.count("ok") == 2
