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


COHORT_WORKER = r'''
import os, sys
sys.path.insert(0, os.environ["GBRS_ROOT"])
import numpy as np, torch.distributed as dist
from gbrs_b200 import reconstruct as rc, synth
from tests import simt_emul

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
tmp = os.environ["GBRS_TMP"]
# the device step is replaced by the emulated kernels (CPU box); everything else is the product's N > 1 path
rc.run_plan_on_device = lambda plan, t, s, device=None, keep_work=False: simt_emul.run_plan_emulated(plan, t, s)
os.environ["GBRS_DATA"] = tmp
files = [os.path.join(tmp, f"s{i}.sample.genes.tpm") for i in range(5)]
os.makedirs(os.path.join(tmp, f"rank{rank}"), exist_ok=True)
outs = [os.path.join(tmp, f"rank{rank}", f"out{i}") for i in range(5)]   # per-rank directory: who wrote what is visible
rc.reconstruct_cohort(files, os.path.join(tmp, "s0.tranprob.npz"), avec_file=os.path.join(tmp, "s0.avecs.npz"),
                      gpos_file=os.path.join(tmp, "s0.ref.gene_pos.ordered.npz"), outbases=outs)
mine = [i for i in range(5) if i % world == rank]
for i in range(5):
    assert os.path.exists(outs[i] + ".genotypes.tsv") == (i in mine)   # a rank handles its own samples only
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok", mine)
'''


def test_two_rank_gloo_reconstruct_cohort(tmp_path):
    """N > 1 path of `reconstruct_cohort`: replicas only -- samples dealt round-robin over the ranks, no collective.
    Two gloo ranks (device step = emulated kernels); together they must produce every sample's files, equal to the
    oracle's."""
    import shutil

    import numpy as np

    if shutil.which("g++") is None:
        pytest.skip("g++ is needed to build the SIMT emulation")
    from gbrs_b200 import synth
    from oracle import reconstruct_oracle as ro
    from tests import reconstruct_checks as chk
    from tests import simt_emul

    simt_emul.build()  # once, before the ranks race for it
    kw = dict(genes_per_chrom=(14, 6), H=4)
    samples = [synth.generate_reconstruct(sample_index=i, **kw) for i in range(5)]
    for i, d in enumerate(samples):
        synth.write_reconstruct_files(d, str(tmp_path), prefix=f"s{i}.")
    script = tmp_path / "worker.py"
    script.write_text(COHORT_WORKER)
    env = dict(os.environ, GBRS_ROOT=ROOT, GBRS_TMP=str(tmp_path), MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="2")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", "29517", str(script)]
    res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert res.stdout.count("ok") == 2
    d0 = samples[0]
    for i, d in enumerate(samples):
        want = ro.reconstruct_tables(d0.chroms, d0.genes, d0.tprob, d0.avecs, d.expr, d0.hname, 1.5, 0.12)
        base = tmp_path / f"rank{i % 2}"
        assert open(base / f"out{i}.genotypes.tsv").read() == chk.tsv_of(want["gtcall"])
        gp = np.load(base / f"out{i}.genoprobs.npz")
        for c in want["gamma"]:
            np.testing.assert_allclose(gp[c], want["gamma"][c], rtol=chk.RTOL, atol=1e-300)


EM_COHORT_WORKER = r'''
import os, sys
sys.path.insert(0, os.environ["GBRS_ROOT"])
import numpy as np, torch.distributed as dist
from gbrs_b200 import cohort, synth
from oracle import em_oracle as eo
from tests import simt_em

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()

class EmulatedEM:
    """EMfactory's surface as far as quantify_cohort uses it, with the device replaced by the emulated library (the real
    kernels and C ABI on the host SIMT shim): this box has no GPU, everything else is the product's cohort path."""
    loaded = []

    def __init__(self, apm, device=None):
        self.apm, self.target_lengths = apm, None
        EmulatedEM.loaded.append(apm.tag)

    def prepare(self, pseudocount=0.0, lenfile=None, read_length=100):
        if self.target_lengths is None:  # the first sample parses the table, the others get it handed over
            assert lenfile == "shared.len"
            self.target_lengths = SHARED_EFF
        self.pat = simt_em.HostPattern(self.apm, gene_of=eo.gene_index(self.apm.num_loci, self.apm.groups))
        self._pattern = self.pat
        self.pat.device = None
        self.pat.prepare(self.target_lengths, pseudocount)

    def run(self, model, tol, max_iters, verbose=False):
        self.out = self.pat.run(model, tol, max_iters)
        self.num_iters = self.out["iters"]

    def get_allelic_expression(self):
        return self.out["theta"]

    def expected_read_counts(self):
        return self.out["counts"]

import torch
torch.cuda.synchronize = lambda *a, **k: None      # no device here
cohort.EMfactory = EmulatedEM
N_SAMPLES = 5
datas = [synth.generate(T=40, N=300, H=4, sample_index=i) for i in range(N_SAMPLES)]
SHARED_EFF = eo.effective_length_table(datas[0].lengths)

def load(i):
    apm = synth.to_apm(datas[i])
    apm.tag = i
    apm.groups = [list(g) for g in apm.groups]      # a fresh, equal list per sample: the cohort must share one
    return apm

stats = {}
local = cohort.quantify_cohort(list(range(N_SAMPLES)), load, model=4, lenfile="shared.len", prefetch=2, stats=stats)
assert sorted(local) == cohort.my_share(N_SAMPLES, rank, world) == EmulatedEM.loaded
assert stats["samples"] == len(local) and stats["nnz_iters"] > 0
merged = cohort.gather_results(local, N_SAMPLES)
assert len(merged) == N_SAMPLES
for i, r in enumerate(merged):                         # every rank holds every sample's result; each equals the oracle's
    d = datas[i]
    oapm = eo.apm_from_pairs(d.T, d.H, d.N, d.pair_class, d.pair_locus, d.pair_mask, d.count)
    o = eo.run(oapm, eo.prepare(oapm, SHARED_EFF, 0.0), 4, SHARED_EFF, None, tol=1e-4)
    assert r["iters"] == o["iters"] and np.abs(r["counts"] - o["counts"]).max() < 1e-9 * np.abs(o["counts"]).max()
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_two_rank_gloo_em_cohort_split(tmp_path):
    """Batched cohort mode (BASELINE config 5) under a 2-rank gloo job: round-robin split, loader thread, shared length
    table and grouping, object all-gather of the per-sample results; the per-sample EM runs through the emulated library
    and must equal the oracle's."""
    import shutil

    if shutil.which("g++") is None:
        pytest.skip("g++ is needed to build the SIMT emulation")
    from tests import simt_em

    simt_em.build()  # once, before two ranks race to build it
    script = tmp_path / "worker.py"
    script.write_text(EM_COHORT_WORKER)
    env = dict(os.environ, GBRS_ROOT=ROOT, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="2")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", "29519", str(script)]
    res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert res.stdout.count("ok") == 2
