"""Host-side pieces of gbrs_b200.emfactory that need no device: the staging of the effective-length table (run on a
stand-in object carrying CPU tensors) and the lifetime of the packer's memory behind the zero-copy array views."""
import gc

import numpy as np
import torch

from gbrs_b200 import synth, utils
from gbrs_b200.emfactory import DevicePattern, PackedPattern


class _Stub:
    """What DevicePattern.set_lengths touches, on the CPU."""
    device = "cpu"

    def __init__(self, T, H):
        self.T, self.H = T, H
        self.efflen = torch.full((T, 8), -7.0, dtype=torch.float64)
        self._stage = torch.full((T, 8), -9.0, dtype=torch.float64)

    def _staging(self):
        return self._stage

    def _sync(self):
        pass


def test_set_lengths_accepts_every_layout():
    T = 50
    for H in (1, 3, 8):
        tl = (np.random.default_rng(H).random((T, H)) + 1).transpose()  # as EMfactory._read_lengths returns it
        for src in (tl, np.ascontiguousarray(tl), tl.tolist(), np.asfortranarray(tl), tl.astype(np.float32)):
            s = _Stub(T, H)
            DevicePattern.set_lengths(s, src)
            want = np.ones((T, 8))
            want[:, :H] = np.asarray(src, dtype=np.float64).T
            assert np.array_equal(s.efflen.numpy(), want)
        s = _Stub(T, H)
        DevicePattern.set_lengths(s, None)
        assert (s.efflen == 1.0).all()
        try:
            DevicePattern.set_lengths(_Stub(T, H), np.ones((H + 1, T)))
            raise AssertionError("a table of the wrong shape must be refused")
        except ValueError:
            pass


def test_packed_arrays_outlive_the_pattern_object():
    d = synth.generate(T=200, N=5000, H=8)
    p = PackedPattern(synth.to_apm(d), gene_of=utils.gene_index(d.T, d.groups()))
    pairs, count = p.arrays["pairs"], p.arrays["count"]
    want = (int(pairs.astype(np.int64).sum()), float(count.sum()))
    assert pairs.base is not None and not pairs.flags.owndata  # a view into the packer's memory, not a copy
    del p
    gc.collect()
    junk = [np.ones(1 << 16) for _ in range(8)]  # churn the allocator
    assert (int(pairs.astype(np.int64).sum()), float(count.sum())) == want and len(junk) == 8
    assert float(count.sum()) == float(d.count.sum())


def test_weighted_matrix_initial_estimate_matches_reference_golden():
    """Stored values other than 1 (and explicit zeros): prepare() takes theta0 from the values exactly like the reference
    (golden written by the unmodified reference, oracle/make_golden_weighted.py).  Host arithmetic only."""
    import os

    from gbrs_b200.emfactory import EMfactory
    from oracle.make_golden_weighted import weighted_values
    from tests import helpers as hp

    z = np.load(os.path.join(hp.GOLDEN, "weighted_m4.npz"))
    d = synth.generate(T=int(z["T"]), N=int(z["N"]), H=int(z["H"]), sample_index=int(z["sample_index"]))
    apm = synth.to_apm(d)
    for h, v in enumerate(weighted_values(apm.data)):
        apm.data[h].data = v
    assert not apm.is_pure_incidence()
    em = EMfactory.__new__(EMfactory)
    em.probability, em.target_lengths = apm, synth.effective_lengths(d)
    assert hp.relerr(em._weighted_theta0(0.0), z["theta0_nopc"]) < 1e-13
    assert hp.relerr(em._weighted_theta0(float(z["pseudocount"])), z["theta0_pc"]) < 1e-13
