"""TEST INFRASTRUCTURE ONLY -- builds tests/simt/hmm_emul.cpp (the HMM kernels of gbrs_b200/csrc/hmm_kernels.cu compiled
for the host SIMT shim) and runs an `HmmPlan` through it with numpy arrays.  Mirrors
gbrs_b200.reconstruct.run_plan_on_device so that the same checks run against the emulated kernels on a CPU box and
against the real ones on the GPU box."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = [os.path.join(HERE, "simt", "hmm_emul.cpp"), os.path.join(HERE, "simt", "simt_shim.h"),
       os.path.join(ROOT, "gbrs_b200", "csrc", "hmm_kernels.cu"), os.path.join(ROOT, "include", "gbrs_em.h")]
OUT_DIR = os.path.join(HERE, "simt", "_build")

EMISSION_ARGS = [C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_double,
                 C.c_void_p, C.c_int32]
RUN_ARGS = [C.c_int32, C.c_void_p, C.c_int32] + [C.c_void_p] * 3 + [C.c_int64] + [C.c_void_p] * 7 + [C.c_int32]


def build(tsan: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    lib = os.path.join(OUT_DIR, "libhmm_emul_tsan.so" if tsan else "libhmm_emul.so")
    if os.path.exists(lib) and all(os.path.getmtime(lib) >= os.path.getmtime(s) for s in SRC):
        return lib
    cmd = ["g++", "-O1", "-g", "-std=c++17", "-shared", "-fPIC", "-pthread", "-ffp-contract=off",
           "-I", os.path.join(ROOT, "include"), "-o", lib + ".tmp", SRC[0]]
    if tsan:
        cmd.insert(1, "-fsanitize=thread")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("g++ failed:\n" + " ".join(cmd) + "\n" + res.stderr)
    os.replace(lib + ".tmp", lib)
    return lib


def bind(path: str):
    lib = C.CDLL(path)
    lib.emul_hmm_emission.restype = C.c_int
    lib.emul_hmm_emission.argtypes = EMISSION_ARGS
    lib.emul_hmm_run.restype = C.c_int
    lib.emul_hmm_run.argtypes = RUN_ARGS
    return lib


_lib = None


def load():
    global _lib
    if _lib is None:
        _lib = bind(build())
    return _lib


def _ptr(a):
    return a.ctypes.data if a.size else None


def run_emission(plan, expr_threshold, sigma, grid=None, lib=None):
    lib = lib or load()
    G, S = plan.expr.shape[0], plan.S
    eprob = np.full((G, S), np.nan)
    expr, aidx = np.ascontiguousarray(plan.expr), np.ascontiguousarray(plan.avec_index)
    avecs, init = np.ascontiguousarray(plan.avecs), np.ascontiguousarray(plan.init)
    grid = grid or max(1, (G + 127) // 128)
    lib.emul_hmm_emission(G, plan.H, _ptr(expr), _ptr(avecs), _ptr(aidx), _ptr(init), float(expr_threshold),
                          float(sigma), _ptr(eprob), grid)
    return eprob


def run_chains(plan, eprob, grid=None, lib=None):
    lib = lib or load()
    G, S = plan.expr.shape[0], plan.S
    chains, init = plan.launch_order(), np.ascontiguousarray(plan.init)
    tprob, eprob = np.ascontiguousarray(plan.tprob), np.ascontiguousarray(eprob)
    out = {k: np.full((G, S), np.nan) for k in ("alpha", "gamma", "delta")}
    out["scaler"] = np.full(G, np.nan)
    backptr = np.full((G, S), 255, dtype=np.uint8)
    states = np.full(plan.n_states_out, -1, dtype=np.int32)
    tlin = np.full(tprob.shape, np.nan)
    lib.emul_hmm_run(len(chains), _ptr(chains), plan.H, _ptr(init), _ptr(eprob), _ptr(tprob), tprob.shape[0], _ptr(tlin),
                     _ptr(out["alpha"]), _ptr(out["scaler"]), _ptr(out["gamma"]), _ptr(out["delta"]), _ptr(backptr),
                     _ptr(states), grid or 2 * len(chains))
    out["states"], out["backptr"], out["eprob"] = states, backptr, eprob
    return out


def run_plan_emulated(plan, expr_threshold, sigma, grid=None):
    return run_chains(plan, run_emission(plan, expr_threshold, sigma), grid=grid)
