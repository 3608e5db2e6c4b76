"""TEST INFRASTRUCTURE ONLY -- golden vector for an alignment matrix with STORED VALUES (weights other than 1, explicit
zeros), produced by the unmodified reference in the build container:  python -m oracle.make_golden_weighted

What the reference does with such a file (and what gbrs_b200 must reproduce): prepare() normalises the stored values per
class and takes theta0 from them (EMfactory.py:95-111); every E-step then starts with probability.reset(), which sets
every stored entry -- explicit zeros included -- to 1 (Sparse3DMatrix.py:220-228).  Writes tests/golden/weighted_m4.npz."""
from __future__ import annotations

import contextlib
import io
import os
import re
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from gbrs_b200 import synth  # noqa: E402
from oracle import ref_harness as rh  # noqa: E402


def weighted_values(mats, seed=7):
    """Deterministic stored values for the H CSC matrices `mats` (as synth.to_csc_list builds them): uniform in (0.25, 4),
    one in eleven an explicit zero -- except that a class whose every stored value came out zero keeps ones (0 / 0 in the
    reference's prepare() is not the case under test)."""
    vals = []
    for h, m in enumerate(mats):
        rng = np.random.default_rng(seed + h)
        v = rng.uniform(0.25, 4.0, m.nnz)
        v[rng.integers(0, 11, m.nnz) == 0] = 0.0
        vals.append(v)
    rows = np.zeros(mats[0].shape[0])
    for m, v in zip(mats, vals):
        np.add.at(rows, m.indices, v)
    for m, v in zip(mats, vals):
        v[rows[m.indices] == 0.0] = 1.0
    return vals


def main():
    d = synth.generate(T=40, N=600, H=4, sample_index=21)
    ref = rh.load_reference()
    rh.install_shim(False)
    apm = rh.build_reference_apm(d)
    for h, v in enumerate(weighted_values(apm.data)):
        apm.data[h].data = v
    out = {}
    for pc in (0.0, 0.7):
        em = ref.EMfactory(apm.copy())
        with tempfile.TemporaryDirectory() as tmp:
            lenfile = os.path.join(tmp, "len.tsv")
            synth.write_length_file(d, lenfile)
            em.prepare(pseudocount=pc, lenfile=lenfile)
        theta0 = np.asarray(em.allelic_expression).copy()
        buf = io.StringIO()
        old = np.geterr()
        try:
            with contextlib.redirect_stdout(buf):
                em.run(model=4, tol=1e-4, max_iters=999, verbose=True)
        finally:
            np.seterr(**old)
        iters = len(re.findall(r"/ 1000000", buf.getvalue()))
        tag = "pc" if pc else "nopc"
        out[f"theta0_{tag}"], out[f"theta_{tag}"], out[f"iters_{tag}"] = theta0, np.asarray(em.allelic_expression).copy(), iters
        out[f"counts_{tag}"] = np.asarray(em.probability.sum(axis=ref.APM.Axis.READ)).copy()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "weighted_m4.npz"), T=d.T, N=d.N, H=d.H, sample_index=21,
                        pseudocount=0.7, **out)
    print({k: (v if np.isscalar(v) or getattr(v, "ndim", 1) == 0 else v.shape) for k, v in out.items()})


if __name__ == "__main__":
    main()
