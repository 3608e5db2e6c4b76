"""TEST INFRASTRUCTURE ONLY -- drives the host-side native code (packer, read-row builder, length parser, table writer,
float formatter) of a sanitizer build through the package's own Python callers.  argv[1]: the sanitized library."""
import ctypes as C
import os
import sys
import tempfile

sys.path.insert(0, os.environ["GBRS_ROOT"])
import numpy as np  # noqa: E402
import scipy.sparse as sp  # noqa: E402

from gbrs_b200 import _lib  # noqa: E402

HOST = ["gbrs_last_error", "gbrs_abi_version", "gbrs_pack_create", "gbrs_pack_get_info", "gbrs_pack_get_array",
        "gbrs_pack_free", "gbrs_write_table", "gbrs_parse_lengths", "gbrs_format_double", "gbrs_rows_create",
        "gbrs_rows_get", "gbrs_rows_free"]
lib = C.CDLL(sys.argv[1])
for n in HOST:
    f = getattr(lib, n)
    f.restype, f.argtypes = _lib.SYMBOLS[n]
_lib._lib = lib  # the package's callers now reach the sanitized host code (no CUDA symbol is touched below)

from gbrs_b200 import synth, utils  # noqa: E402
from gbrs_b200.apm import AlignmentPropertyMatrix as APM  # noqa: E402
from gbrs_b200.compress import read_rows  # noqa: E402
from gbrs_b200.emfactory import EMfactory, PackedPattern  # noqa: E402
from gbrs_b200.quantify import hapmask_bytes  # noqa: E402

CASES = ((dict(T=300, N=20000, H=8), (0, 1), 0, False),
         (dict(T=200, N=9000, H=8, wide_frac=0.05), (1, 3), 8, False),
         (dict(T=200, N=9000, H=8, wide_frac=0.05), (2, 3), 0, False),
         (dict(T=300, N=12000, H=8, with_genotype=True), (0, 1), 0, True),
         (dict(T=150, N=6000, H=3, sample_index=2), (0, 2), 0, False),
         (dict(T=5, N=3, H=1), (0, 1), 0, False),
         (dict(T=50, N=40, H=2), (3, 4), 0, False))
for kw, shard, item_len, mask in CASES:
    d = synth.generate(**kw)
    hm = hapmask_bytes(synth.genotype_mask(d)) if mask else None
    apm = synth.to_apm(d)
    p = PackedPattern(apm, gene_of=utils.gene_index(d.T, d.groups()), hapmask=hm, shard_rank=shard[0],
                      shard_count=shard[1], item_len=item_len)
    q = PackedPattern(apm)  # no gene table
    assert sum(int(np.ascontiguousarray(a).view(np.uint8).sum()) for a in p.arrays.values()) >= 0  # touch every byte
    read_rows(synth.to_csc_list(d), d.T, d.H)
    with tempfile.TemporaryDirectory() as tmp:
        lf = os.path.join(tmp, "len.tsv")
        synth.write_length_file(d, lf)
        em = EMfactory.__new__(EMfactory)
        em.probability = apm
        assert EMfactory._read_lengths(em, lf, 100).shape == (d.H, d.T)
        with open(os.path.join(tmp, "t.tsv"), "w") as fh:
            utils.write_table_rows(fh, apm.lname, np.random.rand(d.H + 1, d.T), notes=None, order=np.arange(d.T)[::-1])
    keep = p.arrays["pairs"]  # a view into the packer's memory must keep it alive (use-after-free otherwise)
    del p, q
    import gc
    gc.collect()
    assert int(keep.astype(np.int64).sum()) >= 0
empty = APM.from_csc([sp.csc_matrix((4, 6)) for _ in range(2)], ["A", "B"], [f"t{i}" for i in range(6)], count=np.ones(4))
assert PackedPattern(empty).info["n_classes"] == 0
buf = C.create_string_buffer(64)
for x in (0.1, 1e22, 5e-324, -0.0, float("inf"), float("nan"), 123456789.125, 1e-7, 2.0 ** 70):
    assert lib.gbrs_format_double(x, buf, 64) > 0
print("HOST-SANITIZER-RUN-OK")
