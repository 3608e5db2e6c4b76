"""TEST INFRASTRUCTURE ONLY -- a numpy walk over the TILE layout (include/gbrs_em.h, built by gbrs_b200/csrc/tile_pack.cpp)
doing, tile by tile, exactly what the fused model-4 kernel does: class weights from the pair planes, per-(locus, bucket)
sums from the tile's locus-major copy, expansion of the nibble buckets to haplotypes, partial sums into slots, and the
per-locus sum over slots.  It checks the LAYOUT (every array the kernel reads) on machines without a GPU."""
from __future__ import annotations

import numpy as np

from gbrs_b200 import _lib


def _u32(blob, off, n):
    return blob[off:off + 4 * n].view(np.uint32)


def tile_views(blob, desc_row):
    a0 = int(desc_row[0]) * 16
    hdr = _u32(blob, a0, _lib.TH_WORDS)
    nc, nl, npl, npairs = (int(hdr[k]) for k in (_lib.TH_CLASSES, _lib.TH_LOCI, _lib.TH_PLANES, _lib.TH_PAIRS))
    ne, ni = int(hdr[_lib.TH_ENTRIES]), int(hdr[_lib.TH_ITEMS])
    assert int(hdr[_lib.TH_A_BYTES]) == int(desc_row[1]) and int(hdr[_lib.TH_B_BYTES]) == int(desc_row[2])
    b0 = a0 + int(hdr[_lib.TH_A_BYTES])
    v = dict(n_classes=nc, n_loci=nl, n_planes=npl, n_pairs=npairs, n_entries=ne, n_items=ni, full=int(hdr[_lib.TH_FLAGS]))
    v["loci"] = _u32(blob, a0 + int(hdr[_lib.TH_OFF_LOCI]), nl)
    v["slots"] = _u32(blob, a0 + int(hdr[_lib.TH_OFF_SLOTS]), nl)
    o = a0 + int(hdr[_lib.TH_OFF_NPLANE])
    v["nplane"] = blob[o:o + 2 * npl].view(np.uint16)
    o = a0 + int(hdr[_lib.TH_OFF_COUNT])
    v["count"] = blob[o:o + 8 * nc].view(np.float64)
    o = a0 + int(hdr[_lib.TH_OFF_PAIRS])
    v["pairs"] = blob[o:o + 2 * (npairs + 3 * npl)].view(np.uint16)
    v["items"] = _u32(blob, b0, ni)
    o = b0 + int(hdr[_lib.TH_OFF_POS])
    v["pos"] = blob[o:o + 2 * ni].view(np.uint16)
    nr = int(hdr[_lib.TH_RUNS])
    o = b0 + int(hdr[_lib.TH_OFF_RUNKEY])
    v["run_key"] = blob[o:o + 2 * nr].view(np.uint16)
    o = b0 + int(hdr[_lib.TH_OFF_RUNFIRST])
    v["run_first"] = blob[o:o + 2 * (nr + 1)].view(np.uint16)
    o = b0 + int(hdr[_lib.TH_OFF_ENTS])
    v["ents"] = blob[o:o + 2 * ne].view(np.uint16)
    return v


def numerator_W(tiled, theta_T8, T, unit=False):
    """W[t][h] = sum over classes hitting (t, h) of count / normaliser, through the tile layout.  theta_T8: [T][8]."""
    blob, desc = tiled.arrays["blob"], tiled.arrays["tile_desc"].reshape(-1, 4)
    partial = np.zeros((tiled.info["n_slots"], 8))
    written = np.zeros(tiled.info["n_slots"], dtype=bool)
    sub = np.zeros((T, 32))
    for m in range(16):
        for b in range(4):
            if (m >> b) & 1:
                sub[:, m] += 1.0 if unit else theta_T8[:, b]
                sub[:, 16 + m] += 1.0 if unit else theta_T8[:, 4 + b]
    for row in desc[: tiled.info["n_tiles"]]:
        v = tile_views(blob, row)
        nc = v["n_classes"]
        tab = sub[v["loci"]]
        s = np.zeros(nc)
        off = 0
        assert v["nplane"].sum() == v["n_pairs"] and (np.diff(v["nplane"].astype(int)) <= 0).all() and v["nplane"][0] == nc
        for p in range(v["n_planes"]):
            n_p = int(v["nplane"][p])
            w = v["pairs"][off:off + n_p].astype(np.int64)
            l, m = w >> 8, w & 255
            assert (l < v["n_loci"]).all() and (m > 0).all()
            s[:n_p] += tab[l, m & 15] + tab[l, 16 + (m >> 4)]
            pad = (n_p + 3) // 4 * 4
            assert (v["pairs"][off + n_p:off + pad] == 0).all()  # padding words add exactly nothing
            off += pad
        wts = v["count"] / s
        acc = np.zeros(v["n_loci"] * 32)
        ni = v["n_items"]
        lens = ((v["items"] >> 16) & 15).astype(int) + 1
        assert sorted(v["pos"].tolist()) == list(range(ni)) and (np.diff(lens) <= 0).all()
        isum = np.zeros(ni)
        covered = np.zeros(v["n_entries"], dtype=int)
        for word, ps, ln in zip(v["items"], v["pos"], lens):
            start = int(word) & 0xFFFF
            assert start + ln <= v["n_entries"]
            covered[start:start + ln] += 1
            idx = v["ents"][start:start + ln]
            assert (idx < nc).all()
            isum[ps] = wts[idx].sum()
        assert (covered == 1).all()
        rf = v["run_first"].astype(int)
        assert rf[0] == 0 and rf[-1] == ni and (np.diff(rf) > 0).all() and (np.diff(v["run_key"].astype(int)) > 0).all()
        for r, key in enumerate(v["run_key"]):
            assert key < v["n_loci"] * 32
            acc[key] = isum[rf[r]:rf[r + 1]].sum()
        acc = acc.reshape(v["n_loci"], 32)
        W = np.zeros((v["n_loci"], 8))
        for h in range(8):
            if (v["full"] >> h) & 1:
                W[:, h] += acc[:, 0]
            half, bit = h >> 2, h & 3
            for val in range(1, 16):
                if (val >> bit) & 1:
                    W[:, h] += acc[:, 16 * half + val]
        assert not written[v["slots"]].any()
        written[v["slots"]] = True
        partial[v["slots"]] = W
    assert written.all()
    out = np.zeros((T, 8))
    ld = tiled.arrays["locus_desc"].reshape(-1, 4)
    seen = np.zeros(T, dtype=bool)
    for t, a, b, _ in ld:
        assert not seen[t]
        seen[t] = True
        out[t] = partial[a:b].sum(axis=0)
    assert seen.all()
    return out
