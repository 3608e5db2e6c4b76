"""The host-side native code -- packer, read-row builder, length parser, table writer (pack.cpp, report.cpp) -- built
with AddressSanitizer + UndefinedBehaviorSanitizer and driven through the package's own Python callers on a spread of
inputs (shards, masks, wide classes, tiny and empty matrices).  The packer works on uninitialised scratch arrays and
hands out views into its own memory; out-of-bounds or misaligned accesses there would corrupt results silently."""
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HERE = os.path.dirname(os.path.abspath(__file__))


def test_host_code_is_clean_under_asan_and_ubsan():
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    rts = [subprocess.run(["g++", f"-print-file-name={n}"], capture_output=True, text=True).stdout.strip()
           for n in ("libasan.so", "libubsan.so")]
    if not all(os.path.isabs(r) and os.path.exists(r) for r in rts):
        pytest.skip("sanitizer runtimes not found")
    out_dir = os.path.join(HERE, "simt", "_build")
    os.makedirs(out_dir, exist_ok=True)
    lib = os.path.join(out_dir, "libgbrs_host_asan.so")
    srcs = [os.path.join(ROOT, "gbrs_b200", "csrc", "pack.cpp"), os.path.join(ROOT, "gbrs_b200", "csrc", "report.cpp"),
            os.path.join(HERE, "sanitize", "host_stub.cpp")]
    if not os.path.exists(lib) or any(os.path.getmtime(s) > os.path.getmtime(lib) for s in srcs):
        cmd = ["g++", "-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-O1", "-g", "-std=c++17", "-shared",
               "-fPIC", "-fopenmp", "-I", os.path.join(ROOT, "include"), "-o", lib] + srcs
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            pytest.skip("sanitizer build failed: " + res.stderr[-500:])
    env = dict(os.environ, GBRS_ROOT=ROOT, LD_PRELOAD=" ".join(rts), OMP_NUM_THREADS="4",
               ASAN_OPTIONS="detect_leaks=0:halt_on_error=0", UBSAN_OPTIONS="print_stacktrace=1")
    res = subprocess.run([sys.executable, os.path.join(HERE, "sanitize", "drive_host.py"), lib], env=env,
                         capture_output=True, text=True, timeout=900)
    if "HOST-SANITIZER-RUN-OK" not in res.stdout:
        if "AddressSanitizer" in res.stderr or "runtime error" in res.stderr:
            pytest.fail(res.stderr[-4000:])
        pytest.skip("the sanitizer run did not complete here: " + res.stderr[-600:])
    assert "AddressSanitizer" not in res.stderr and "runtime error" not in res.stderr, res.stderr[-4000:]
