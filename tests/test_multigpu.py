"""Two real GPUs (skipped on a single-GPU box): `EMfactory(shard=True)` under torchrun / NCCL.  Every rank packs its
own contiguous slice of the alignment classes, the T x 8 numerator is summed over the ranks once per update (push: the
one-launch push form over NVLink peer memory, the default; pull: the owner loads its slice from every peer; nvls: pull with
the in-switch reduction where the box offers a multicast mapping; nccl: plain all-reduce), and all ranks must reach the
reference's iteration count and results."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
sys.path.insert(0, os.environ["GBRS_ROOT"])
import numpy as np, torch, torch.distributed as dist
from gbrs_b200 import synth
from gbrs_b200.emfactory import EMfactory
from tests import helpers as hp

local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
for name in ("em_small_m4", "em_small_m2", "em_small_m1_biggenes", "em_small_m3_diploid", "c1_model4"):
    g = hp.load_golden(name)
    d = hp.synth_from_golden(g)
    apm = synth.to_apm(d)
    if g["masked"]:
        apm.multiply(g["gtmask"], axis=2); apm.eliminate_zeros()
    em = EMfactory(apm, shard=True, poll_every=3)
    want_fused = os.environ.get("GBRS_XCHG", "push") in ("push", "tag", "fused", "pull", "nvls", "p2p")
    em.target_lengths = synth.effective_lengths(d)
    em.prepare(pseudocount=g["pseudocount"])
    assert hp.relerr(em.get_allelic_expression(), g["theta0"]) < 1e-9
    em.run(model=g["model"], tol=g["tol"], max_iters=g["max_iters"], verbose=False)
    assert em.num_iters == g["iters"], (name, em.num_iters, g["iters"])
    if want_fused and rank == 0 and name == "em_small_m4":
        print("fused exchange in use:", em.fused_exchange, "mode:", em.exchange_mode, "NVLS:", em.nvls_exchange)
        assert em.fused_exchange and em.exchange_mode == {"fused": "push", "p2p": "pull"}.get(os.environ["GBRS_XCHG"], os.environ["GBRS_XCHG"])
        assert not (os.environ.get("GBRS_XCHG") == "pull" and em.nvls_exchange)  # (push broadcasts through the multicast mapping where there is one)
    assert hp.relerr(em.allelic_expression, g["theta"]) < 1e-9
    assert hp.relerr(em.expected_read_counts(), g["counts"]) < 1e-9
    np.testing.assert_allclose(em.err_history, g["errs"], rtol=1e-7, atol=1e-7)
    # every rank holds bit-identical results
    t = torch.from_numpy(em.allelic_expression.copy()).cuda()
    lo, hi = t.clone(), t.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    assert torch.equal(lo, hi)
dist.destroy_process_group()
print("rank", rank, "ok")
'''


@pytest.mark.parametrize("xchg", ["push", "tag", "pull", "nvls", "nccl"])
def test_two_gpu_sharded_run_matches_reference(tmp_path, xchg):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, GBRS_ROOT=ROOT, MASTER_ADDR="127.0.0.1", GBRS_XCHG=xchg)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", "29517", str(script)]
    res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=240)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert res.stdout.count("ok") == 2
    if xchg != "nccl":
        assert "fused exchange in use: True" in res.stdout, res.stdout[-2000:]
        print(res.stdout[-400:])
