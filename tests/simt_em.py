"""TEST INFRASTRUCTURE ONLY -- the EM library (gbrs_b200/csrc/em_kernels.cu: kernels AND launchers AND the C ABI) built
for the host SIMT shim, so that the CPU suite can run the code the GPU runs against the reference's golden vectors.

The product source is not modified: a copy is rewritten mechanically at test time --
  * `#include <cuda_runtime.h>`            -> the fake runtime (tests/simt/cuda_runtime_fake.h: device memory = host memory,
                                              synchronous streams, no graph capture, tiny occupancy)
  * `kernel<<<grid, block, 0, s>>>(args);` -> `SIMT_LAUNCH(grid, block, kernel(args));`  (one OS thread per CUDA thread)
  * `asm volatile("<ptx>" ...);`           -> C++ atomics / plain accesses for the system-scope flag and numerator accesses
    of the multi-GPU exchange (ranks = library instances running concurrently on shared host memory), a trap for multimem;
    the `rcp.approx.ftz.f64` seed of fast_div -> a 20-bit reciprocal
and compiled with g++ together with the host sources (pack.cpp, report.cpp).  The result exports the same `gbrs_*`
symbols as libgbrs_em.so; `HostPattern` below is DevicePattern with numpy arrays instead of torch CUDA tensors."""
from __future__ import annotations

import ctypes as C
import os
import re
import subprocess

import numpy as np

from gbrs_b200 import _lib

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(ROOT, "gbrs_b200", "csrc")
OUT = os.path.join(HERE, "simt", "_build")

LAUNCH = re.compile(r"(\bk_[a-z0-9_]+(?:<[^;<>]*>)?)\s*<<<(.+?),\s*(k[A-Za-z]*Threads|32),\s*(?:0|[A-Za-z_][\w.]*),\s*(.+?)>>>\(\s*(.*?)\);", re.S)
ASM = re.compile(r"asm volatile\(.*?\);", re.S)
RCP = re.compile(r'asm\("rcp\.approx\.ftz\.f64 %0, %1;" : "=d"\((\w+)\) : "d"\((\w+)\)\);')


PTX = [  # the system-scope accesses of the multi-GPU exchange -> host equivalents (see cuda_runtime_fake.h)
    (r'asm volatile\("mov\.u64 %0, %globaltimer;" : "=l"\((\w+)\)\);', r"\1 = simt_globaltimer_ns();"),
    (r'asm volatile\("st\.relaxed\.sys\.global\.f64 \[%0\], %1;" ::"l"\((\w+)\), "d"\((\w+)\) : "memory"\);', r"*\1 = \2;"),
    (r'asm volatile\("ld\.acquire\.sys\.global\.u32 %0, \[%1\];" : "=r"\((\w+)\) : "l"\((\w+)\) : "memory"\);',
     r"\1 = simt_ld_acquire_u32(\2);"),
    (r'asm volatile\("st\.relaxed\.sys\.global\.u32 \[%0\], %1;" ::"l"\((\w+)\), "r"\((\w+)\) : "memory"\);',
     r"simt_st_release_u32(\1, \2);"),
    (r'asm volatile\("ld\.relaxed\.sys\.global\.v2\.f64 \{%0, %1\}, \[%2\];" : "=d"\((\w+)\.x\), "=d"\(\w+\.y\) : "l"\((\w+)\) : "memory"\);',
     r"\1.x = \2[0]; \1.y = \2[1];"),
    (r'asm volatile\("st\.relaxed\.sys\.global\.v2\.f64 \[%0\], \{%1, %2\};" ::"l"\((\w+)\), "d"\((\w+)\.x\), "d"\(\w+\.y\) : "memory"\);',
     r"\1[0] = \2.x; \1[1] = \2.y;"),
    # tag form: self-validating doubles (the sign bit says which exchange wrote them) -> 8-byte atomics
    (r'asm volatile\("ld\.relaxed\.sys\.global\.b64 %0, \[%1\];" : "=l"\((\w+)\) : "l"\((\w+)\) : "memory"\);',
     r"\1 = simt_ld_acquire_u64(\2);"),
    (r'asm volatile\("st\.relaxed\.sys\.global\.b64 \[%0\], %1;" ::"l"\((\w+)\), "l"\((\w+)\) : "memory"\);',
     r"simt_st_release_u64(\1, \2);"),
    (r'asm volatile\("ld\.relaxed\.sys\.global\.v2\.b64 \{%0, %1\}, \[%2\];" : "=l"\((\w+)\.x\), "=l"\(\w+\.y\) : "l"\((\w+)\) : "memory"\);',
     r"\1.x = simt_ld_acquire_u64(\2); \1.y = simt_ld_acquire_u64(\2 + 1);"),
    (r'asm volatile\("st\.relaxed\.sys\.global\.v2\.b64 \[%0\], \{%1, %2\};" ::"l"\((\w+)\), "l"\((\w+)\.x\), "l"\(\w+\.y\) : "memory"\);',
     r"simt_st_release_u64(\1, \2.x); simt_st_release_u64(\1 + 1, \2.y);"),
]


def transform(src: str) -> tuple:
    n_launch = len(LAUNCH.findall(src))
    out = LAUNCH.sub(lambda m: f"SIMT_LAUNCH({m.group(2)}, {m.group(3)}, {m.group(1)}({m.group(5)}));", src)
    n_asm = 0
    for pat, rep in PTX:
        out, k = re.subn(pat, rep, out)
        assert k == 1, pat
        n_asm += k
    n_asm += len(ASM.findall(out))
    out = ASM.sub("SIMT_PTX_UNSUPPORTED();", out)
    out, n_rcp = RCP.subn(r"\1 = simt_rcp_approx(\2);", out)
    assert n_rcp == 1 and "asm(" not in out and "asm volatile" not in out
    assert "#include <cuda_runtime.h>" in out
    out = out.replace("#include <cuda_runtime.h>", '#include "cuda_runtime_fake.h"')
    assert "<<<" not in out, "a kernel launch was not rewritten"
    return out, n_launch, n_asm


def build_ec() -> str:
    """ec_kernels.cu (`gbrs compress`: hashing, exact grouping, first-appearance numbering + launcher + C ABI) for the host
    shim; the two CUB library calls (radix sort, prefix sum) are served by tests/simt/fake_cub (a stable sort, a loop)."""
    os.makedirs(OUT, exist_ok=True)
    lib = os.path.join(OUT, "libgbrs_ec_emul.so")
    src = os.path.join(CSRC, "ec_kernels.cu")
    deps = [src, os.path.abspath(__file__)] + [os.path.join(HERE, "simt", n) for n in (
        "simt_shim.h", "cuda_runtime_fake.h", "fake_cub/cub/device/device_radix_sort.cuh", "fake_cub/cub/device/device_scan.cuh")]
    if os.path.exists(lib) and all(os.path.getmtime(lib) >= os.path.getmtime(d) for d in deps):
        return lib
    text = open(src).read()
    n_launch = len(LAUNCH.findall(text))
    text = LAUNCH.sub(lambda m: f"SIMT_LAUNCH({m.group(2)}, {m.group(3)}, {m.group(1)}({m.group(5)}));", text)
    assert n_launch >= 5 and "<<<" not in text
    text = text.replace("#include <cuda_runtime.h>", '#include "cuda_runtime_fake.h"')
    gen = os.path.join(OUT, "ec_kernels_emul.cpp")
    with open(gen, "w") as fh:
        fh.write("// GENERATED by tests/simt_em.py from gbrs_b200/csrc/ec_kernels.cu -- do not edit\n" + text +
                 "\n#include <string>\nvoid gbrs_set_error(const std::string&) {}\n")
    cmd = ["g++", "-O1", "-g", "-std=c++17", "-shared", "-fPIC", "-pthread", "-Wno-unknown-pragmas",
           "-I", os.path.join(HERE, "simt", "fake_cub"), "-I", os.path.join(HERE, "simt"),
           "-I", os.path.join(ROOT, "include"), "-o", lib + ".tmp", gen]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("g++ failed:\n" + res.stderr[-6000:])
    os.replace(lib + ".tmp", lib)
    return lib


def build_gpack() -> str:
    """gpu_pack.cu (the device-side packer: kernels + launcher + C ABI) for the host shim, CUB served by tests/simt/fake_cub."""
    os.makedirs(OUT, exist_ok=True)
    lib = os.path.join(OUT, "libgbrs_gpack_emul.so")
    src = os.path.join(CSRC, "gpu_pack.cu")
    deps = [src, os.path.abspath(__file__), os.path.join(ROOT, "include", "gbrs_em.h")] + [os.path.join(HERE, "simt", n) for n in (
        "simt_shim.h", "cuda_runtime_fake.h", "fake_cub/cub/device/device_radix_sort.cuh", "fake_cub/cub/device/device_scan.cuh")]
    if os.path.exists(lib) and all(os.path.getmtime(lib) >= os.path.getmtime(d) for d in deps):
        return lib
    text = open(src).read()
    n_launch = len(LAUNCH.findall(text))
    text = LAUNCH.sub(lambda m: f"SIMT_LAUNCH({m.group(2)}, {m.group(3)}, {m.group(1)}({m.group(5)}));", text)
    assert n_launch >= 18 and "<<<" not in text, n_launch
    text = text.replace("#include <cuda_runtime.h>", '#include "cuda_runtime_fake.h"')
    gen = os.path.join(OUT, "gpu_pack_emul.cpp")
    with open(gen, "w") as fh:
        fh.write("// GENERATED by tests/simt_em.py from gbrs_b200/csrc/gpu_pack.cu -- do not edit\n" + text +
                 "\n#include <string>\nstatic std::string g_err;\nvoid gbrs_set_error(const std::string& s) { g_err = s; }\n"
                 'extern "C" const char* gbrs_last_error(void) { return g_err.c_str(); }\n')
    cmd = ["g++", "-O1", "-g", "-std=c++17", "-shared", "-fPIC", "-pthread", "-Wno-unknown-pragmas",
           "-I", os.path.join(HERE, "simt", "fake_cub"), "-I", os.path.join(HERE, "simt"),
           "-I", os.path.join(ROOT, "include"), "-o", lib + ".tmp", gen]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("g++ failed:\n" + res.stderr[-6000:])
    os.replace(lib + ".tmp", lib)
    return lib


def emulated_device_pack(apm, gene_of=None, hapmask=None, shard_rank=0, shard_count=1, item_len=0):
    """gbrs_pack_device through the host shim: returns (info dict, {array name: numpy array}) -- "device" memory is host
    memory handed out by a numpy-backed allocator."""
    from gbrs_b200.emfactory import pack_input

    lib = C.CDLL(build_gpack())
    lib.gbrs_pack_device.restype, lib.gbrs_pack_device.argtypes = _lib.SYMBOLS["gbrs_pack_device"]
    lib.gbrs_last_error.restype = C.c_char_p
    inp, keep = pack_input(apm, gene_of=gene_of, hapmask=hapmask, shard_rank=shard_rank, shard_count=shard_count, item_len=item_len)
    bufs = {}

    def alloc(nbytes, tag, _user):
        a = np.zeros(max(int(nbytes), 1) + 64, dtype=np.uint8)
        bufs.setdefault(tag.decode(), []).append((a, int(nbytes)))
        return a.ctypes.data

    cb = _lib.ALLOC_FN(alloc)
    info, out = _lib.PackInfo(), _lib.DevicePack()
    rc = lib.gbrs_pack_device(C.byref(inp), cb, None, None, C.byref(info), C.byref(out))
    if rc != 0:
        msg = (lib.gbrs_last_error() or b"").decode()
        if rc == _lib.GBRS_E_LIMIT:
            raise NotImplementedError(msg)
        raise EmulError(f"rc={rc}: {msg}")
    del keep
    inf = {f: getattr(info, f) for f, _ in _lib.PackInfo._fields_}
    inf["bucket_class0"], inf["bucket_pair0"] = list(info.bucket_class0), list(info.bucket_pair0)
    entry_t = np.uint32 if info.entry_bytes == 4 else np.uint64
    dts = {"rowptr": np.uint32, "pairs": np.uint32, "count": np.float64, "runptr": np.uint32, "ent_cls": entry_t, "ent_pair": entry_t,
           "ent_run": entry_t, "item_desc": np.uint32, "locus_desc": np.uint32, "gene_of": np.int32, "gene_ptr": np.uint32,
           "gene_loci": np.uint32}
    arrays = {}
    for name, dt in dts.items():
        a, nbytes = bufs[name][-1]
        assert getattr(out, name) == a.ctypes.data, name
        arrays[name] = a[:nbytes].view(dt).copy()
    return inf, arrays


def emulated_equivalence_classes(rowptr, words, count=None, device=None):
    """gbrs_b200.compress.equivalence_classes with host arrays and the emulated `gbrs_ec_build`."""
    lib = C.CDLL(build_ec())
    for name in ("gbrs_ec_workspace_bytes", "gbrs_ec_build"):
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = _lib.SYMBOLS[name]
    n = int(rowptr.shape[0]) - 1
    if n == 0:
        return np.zeros(0, np.uint32), np.zeros(0, np.uint32), np.zeros(0)
    rp = np.ascontiguousarray(rowptr, dtype=np.uint32)
    w = np.ascontiguousarray(words, dtype=np.uint32) if len(words) else np.zeros(4, np.uint32)
    cnt = None if count is None else np.ascontiguousarray(count, dtype=np.float64)
    cls, first, ccount = np.zeros(n, np.uint32), np.zeros(n, np.uint32), np.zeros(n)
    nbytes = C.c_int64()
    assert lib.gbrs_ec_workspace_bytes(n, C.byref(nbytes)) == 0
    work = np.zeros(nbytes.value, np.uint8)
    n_classes, collisions = C.c_int64(), C.c_int64()
    for seed in (0x243F6A8885A308D3, 0x13198A2E03707344):
        rc = lib.gbrs_ec_build(n, rp.ctypes.data, w.ctypes.data, None if cnt is None else cnt.ctypes.data, seed,
                               cls.ctypes.data, first.ctypes.data, ccount.ctypes.data, work.ctypes.data, nbytes.value,
                               None, C.byref(n_classes), C.byref(collisions))
        assert rc == 0, rc
        if collisions.value == 0:
            break
    k = int(n_classes.value)
    return cls, first[:k].copy(), ccount[:k].copy()


def build(tsan: bool = False) -> str:
    os.makedirs(OUT, exist_ok=True)
    lib = os.path.join(OUT, "libgbrs_em_emul_tsan.so" if tsan else "libgbrs_em_emul.so")
    # the ThreadSanitizer build leaves the OpenMP host sources out (libgomp is not instrumented: false reports)
    srcs = [os.path.join(CSRC, n) for n in (("em_kernels.cu",) if tsan else ("em_kernels.cu", "pack.cpp", "tile_pack.cpp", "report.cpp"))]
    deps = srcs + [os.path.join(HERE, "simt", n) for n in ("simt_shim.h", "cuda_runtime_fake.h")] + [
        os.path.join(ROOT, "include", "gbrs_em.h"), os.path.abspath(__file__)]
    if os.path.exists(lib) and all(os.path.getmtime(lib) >= os.path.getmtime(d) for d in deps):
        return lib
    text, n_launch, n_asm = transform(open(srcs[0]).read())
    assert n_launch >= 30 and n_asm >= 4, (n_launch, n_asm)
    gen = os.path.join(OUT, "em_kernels_emul_tsan.cpp" if tsan else "em_kernels_emul.cpp")
    with open(gen, "w") as fh:
        fh.write("// GENERATED by tests/simt_em.py from gbrs_b200/csrc/em_kernels.cu -- do not edit\n" + text)
    cmd = ["g++", "-O1", "-g", "-std=c++17", "-shared", "-fPIC", "-pthread", "-fopenmp", "-ffp-contract=off",
           "-Wno-unknown-pragmas", "-I", os.path.join(HERE, "simt"), "-I", os.path.join(ROOT, "include"),
           "-o", lib + ".tmp", gen] + srcs[1:]
    if tsan:
        cmd.insert(1, "-fsanitize=thread")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("g++ failed:\n" + res.stderr[-6000:])
    os.replace(lib + ".tmp", lib)
    return lib


EM_SYMBOLS = ["gbrs_last_error", "gbrs_abi_version", "gbrs_em_prepare_local", "gbrs_em_prepare_finish", "gbrs_em_set_theta",
              "gbrs_em_run_begin", "gbrs_em_launch_local", "gbrs_em_launch_estep", "gbrs_em_launch_update", "gbrs_em_run", "gbrs_em_read_ctrl",
              "gbrs_em_alignment_counts"]
_emul = {}


def load(tsan: bool = False):
    if tsan not in _emul:
        lib = C.CDLL(build(tsan))
        for name in EM_SYMBOLS:
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = _lib.SYMBOLS[name]
        assert lib.gbrs_abi_version() == _lib.ABI_VERSION
        _emul[tsan] = lib
    return _emul[tsan]


def load_instance(tag: str, tsan: bool = False):
    """A private copy of the emulated library (own static state: grid / block indices, barriers, shared memory), so that
    several `ranks` can run kernels at the same time in different host threads."""
    import shutil

    src = build(tsan)
    dst = os.path.join(OUT, f"libgbrs_em_emul_{'tsan_' if tsan else ''}{tag}.so")
    if not os.path.exists(dst) or os.path.getmtime(dst) < os.path.getmtime(src):
        shutil.copyfile(src, dst)
    lib = C.CDLL(dst)
    for name in EM_SYMBOLS:
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = _lib.SYMBOLS[name]
    return lib


class EmulError(RuntimeError):
    pass


class HostPattern:
    """gbrs_b200.emfactory.DevicePattern with numpy arrays: the packed incidence (from the product's own packer), the
    state vectors and the `gbrs_em_dev` descriptor pointing at them; the C ABI calls go to the emulated library."""

    def __init__(self, apm, gene_of=None, hapmask=None, item_len=0, lib=None, shard_rank=0, shard_count=1, tiles=None):
        """`tiles`: dict of tile caps (or {}) to run model 4 / prepare through the fused tile kernel; None = two-pass."""
        from gbrs_b200.emfactory import ERR_LOG_CAP, PackedPattern, TiledPattern

        self.lib = lib or load()
        self.packed = PackedPattern(apm, gene_of=gene_of, hapmask=hapmask, item_len=item_len, shard_rank=shard_rank,
                                    shard_count=shard_count)
        self.info, self.T, self.H = self.packed.info, self.packed.T, self.packed.H
        i = self.info
        self.arr = {k: (np.ascontiguousarray(a) if a.size else np.zeros(16, np.uint8)) for k, a in self.packed.arrays.items()}
        nw = max(i["n_classes"], i["n_pairs"], 8 * i["n_runs"] if self.packed.has_genes else 0, 1) + 8
        T = self.T
        z = np.zeros
        self.theta, self.efflen, self.acc, self.iso = z((2, T, 8)), np.ones((T, 8)), z((T, 8)), z((2, T))
        self.weights, self.subsets, self.wit = z(nw), z((T, 32)), z((max(i["n_items"], 1), 8))
        self.part, self.gene_hap, self.gamma = z(_lib.GBRS_PART_SLOTS), z((max(i["n_gene_ids"], 1), 8)), z(T)
        self.err_log, self.scal, self.ctrl = z(ERR_LOG_CAP), z(8), np.zeros(16, np.int32)
        d = _lib.EmDev()
        d.T, d.H, d.n_gene_ids, d.entry_bytes = T, self.H, i["n_gene_ids"], i["entry_bytes"]
        d.n_classes, d.n_pairs, d.n_runs, d.n_items = i["n_classes"], i["n_pairs"], i["n_runs"], i["n_items"]
        d.n_long_items, d.n_entries, d.n_ranks, d.max_iters_cap = i["n_long_items"], i["n_entries"], shard_count, ERR_LOG_CAP
        d.n_deep_loci = i["n_deep_loci"]
        for k in range(_lib.GBRS_KMAX + 2):
            d.bucket_class0[k], d.bucket_pair0[k] = i["bucket_class0"][k], i["bucket_pair0"][k]
        for k in ("rowptr", "pairs", "count", "runptr", "ent_cls", "ent_pair", "ent_run", "item_off", "item_order",
                  "item_desc", "locus_order", "locus_desc", "locus_item_ptr"):
            setattr(d, k, self.arr[k].ctypes.data)
        if self.packed.has_genes:
            for k in ("gene_of", "gene_ptr", "gene_loci"):
                setattr(d, k, self.arr[k].ctypes.data)
            d.gene_hap, d.gamma = self.gene_hap.ctypes.data, self.gamma.ctypes.data
        for k in ("theta", "efflen", "acc", "iso", "weights", "subsets", "wit", "part", "err_log", "scal", "ctrl"):
            setattr(d, k, getattr(self, k).ctypes.data)
        self.tiled = None
        if tiles is not None:
            self.tiled = TiledPattern(self.packed, **tiles)
            ti = self.tiled.info
            self.tarr = {k: (np.ascontiguousarray(a) if a.size else np.zeros(16, a.dtype)) for k, a in self.tiled.arrays.items()}
            self.tile_partial = z((max(ti["n_slots"], 1), 8))
            d.tile_blob, d.tile_desc = self.tarr["blob"].ctypes.data, self.tarr["tile_desc"].ctypes.data
            d.tile_locus_desc, d.tile_partial = self.tarr["locus_desc"].ctypes.data, self.tile_partial.ctypes.data
            d.n_tiles, d.n_tile_slots = ti["n_tiles"], ti["n_slots"]
            d.tile_max_classes, d.tile_max_loci, d.tile_max_items = ti["max_classes"], ti["max_loci"], ti["max_items"]
            d.tile_max_a_bytes, d.tile_max_b_bytes = ti["max_part_a_bytes"], ti["max_part_b_bytes"]
            d.tile_n_deep_loci = ti["n_deep_loci"]
        self.desc = d

    def check(self, rc):
        if rc != 0:
            msg = (self.lib.gbrs_last_error() or b"").decode()
            if rc == _lib.GBRS_E_NUMERIC:
                raise FloatingPointError(msg)
            raise EmulError(f"rc={rc}: {msg}")

    def current_theta(self):
        return self.theta[int(self.ctrl[_lib.CTRL_PARITY])][:, : self.H].T.copy()

    def prepare(self, eff_HT=None, pseudocount=0.0):
        self.efflen[:] = 1.0
        if eff_HT is not None:
            self.efflen[:, : self.H] = np.asarray(eff_HT).T
        self.check(self.lib.gbrs_em_prepare_local(C.byref(self.desc), None))
        self.check(self.lib.gbrs_em_prepare_finish(C.byref(self.desc), float(pseudocount), None))
        return self.current_theta()

    def run(self, model, tol=1e-4, max_iters=999, poll_every=4):
        iters, errs = C.c_int32(0), np.zeros(max(max_iters, 1))
        self.check(self.lib.gbrs_em_run(C.byref(self.desc), int(model), float(tol), int(max_iters), int(poll_every), None,
                                        C.byref(iters), errs.ctypes.data))
        n = int(iters.value)
        return dict(theta=self.current_theta(), counts=self.acc[:, : self.H].T.copy(), iters=n, errs=errs[:n].copy())

    def alignment_counts(self, gene_level=False, n_real_genes=0):
        rows = n_real_genes if gene_level else self.T
        alloc = max(self.info["n_gene_ids"], rows) if gene_level else self.T
        aln, uniq, lu = np.zeros((alloc, 8)), np.zeros((alloc, 8)), np.zeros(alloc)
        self.check(self.lib.gbrs_em_alignment_counts(C.byref(self.desc), int(gene_level), int(n_real_genes),
                                                     aln.ctypes.data, uniq.ctypes.data, lu.ctypes.data, None))
        return aln[:rows, : self.H].T.copy(), uniq[:rows, : self.H].T.copy(), lu[:rows].copy()
