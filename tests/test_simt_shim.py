"""Self-test of the host SIMT shim (tests/simt/simt_shim.h), the tool that lets the CPU suite run kernel source without
a GPU: small plain-CUDA kernels that use every primitive the shim provides (warp shuffles with and without width, ballot,
votes, warp reductions, integer and double atomics on shared and global memory, vector loads, bit helpers) against
numpy.  The same kernel file is also compiled with nvcc for sm_100a, so the shim cannot drift from CUDA's signatures."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SIMT = os.path.join(HERE, "simt")
pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="g++ is needed to build the SIMT emulation")


@pytest.fixture(scope="module")
def lib():
    out = os.path.join(SIMT, "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libselftest.so")
    srcs = [os.path.join(SIMT, n) for n in ("selftest_emul.cpp", "selftest_kernels.cu", "simt_shim.h")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        res = subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-shared", "-fPIC", "-pthread", "-o", so, srcs[0]],
                             capture_output=True, text=True)
        assert res.returncode == 0, res.stderr
    return C.CDLL(so)


def p(a):
    return C.c_void_p(a.ctypes.data)


@pytest.mark.parametrize("grid,block", [(3, 128), (2, 96), (1, 32)])  # four warps, three warps, one warp
def test_block_sum_shuffles_and_double_atomics(lib, grid, block):
    x = np.random.default_rng(grid * block).random(5000)
    total, per_block = np.zeros(1), np.zeros(grid)
    lib.emul_block_sum(p(x), C.c_int64(x.size), p(total), p(per_block), grid, block)
    idx = np.arange(x.size)
    want = np.array([x[(idx // block) % grid == b].sum() for b in range(grid)])
    np.testing.assert_allclose(per_block, want, rtol=1e-13)
    np.testing.assert_allclose(total[0], x.sum(), rtol=1e-13)


def test_ballot_compaction_and_integer_atomics(lib):
    x = np.random.default_rng(1).integers(-50, 50, 3000).astype(np.int32)
    out, counter = np.zeros(x.size, np.int32), np.zeros(1, np.uint32)
    lib.emul_compact_positive(p(x), C.c_int32(x.size), p(out), p(counter), 2, 128)
    k = int(counter[0])
    assert k == int((x > 0).sum())
    assert sorted(out[:k]) == sorted(x[x > 0])  # order between warps is free, the multiset is not


def test_group_scan_votes_and_warp_reductions(lib):
    x = np.random.default_rng(2).integers(0, 100, 128).astype(np.uint32)
    x[64:96] = np.minimum(x[64:96], 80)  # a warp without values above 90
    scan8, wmax, votes = np.zeros(128, np.uint32), np.zeros(128, np.uint32), np.zeros(128, np.int32)
    lib.emul_group_scan(p(x), p(scan8), p(wmax), p(votes), 2, 64)
    assert np.array_equal(scan8, np.cumsum(x.reshape(-1, 8), axis=1).reshape(-1).astype(np.uint32))
    assert np.array_equal(wmax, np.repeat(x.reshape(-1, 32).max(axis=1), 32))
    want = np.repeat([(1 if (w > 90).any() else 0) | (2 if (w < 100).all() else 0) for w in x.reshape(-1, 32)], 32)
    assert np.array_equal(votes, want) and set(votes[64:96]) == {2}


def test_shared_and_global_double_atomics_vector_loads(lib):
    rng = np.random.default_rng(3)
    xy = np.stack((rng.integers(0, 1000, 4000).astype(np.float64), rng.integers(1, 9, 4000).astype(np.float64)), axis=1)
    hist, checksum = np.zeros(16), np.zeros(1, np.uint64)
    lib.emul_histogram(p(np.ascontiguousarray(xy)), C.c_int32(4000), p(hist), p(checksum), 3, 64)
    bins = xy[:, 0].astype(np.int64) & 15
    assert np.array_equal(hist, np.bincount(bins, weights=xy[:, 1], minlength=16))  # integer-valued sums: exact
    assert int(checksum[0]) == int(sum(bin(int(b)).count("1") for b in bins))


@pytest.mark.skipif(shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"), reason="no nvcc")
def test_selftest_kernels_are_valid_cuda(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    res = subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-c", os.path.join(SIMT, "selftest_kernels.cu"),
                          "-o", str(tmp_path / "selftest.o")], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
