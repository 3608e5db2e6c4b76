"""TEST INFRASTRUCTURE ONLY -- a numpy walk over the TILE layout (include/gbrs_em.h, built by gbrs_b200/csrc/tile_pack.cpp)
doing, tile by tile, exactly what the fused model-4 kernel does: class weights from the pair planes, item sums from the
sliced-ELL entry words, bucket sums per key run, expansion of the nibble buckets to haplotypes, partial sums into slots,
and the per-locus sum over slots.  It checks the LAYOUT (every array the kernel reads) on machines without a GPU."""
from __future__ import annotations

import numpy as np

from gbrs_b200 import _lib


def _al16(x):
    return (x + 15) // 16 * 16


def tile_views(blob, desc_row):
    a0 = int(desc_row[0]) * 16
    nc, nl = int(desc_row[1]) & 0xFFFF, int(desc_row[1]) >> 16
    npl, nr = int(desc_row[2]) & 0xFFFF, int(desc_row[2]) >> 16
    ni, ns = int(desc_row[3]) & 0xFFFF, int(desc_row[3]) >> 16
    a_bytes, full = int(desc_row[4]), int(desc_row[5])
    hdr = blob[a0:a0 + 4 * _lib.TH_WORDS].view(np.uint32)
    assert [int(hdr[k]) for k in (_lib.TH_CLASSES, _lib.TH_LOCI, _lib.TH_PLANES, _lib.TH_RUNS, _lib.TH_ITEMS,
                                  _lib.TH_SLICES, _lib.TH_A_BYTES, _lib.TH_FULL)] == [nc, nl, npl, nr, ni, ns, a_bytes, full]
    assert a_bytes + int(hdr[_lib.TH_B_BYTES]) == int(desc_row[7])
    v = dict(n_classes=nc, n_loci=nl, n_planes=npl, n_runs=nr, n_items=ni, n_slices=ns, full=full,
             n_pairs=int(hdr[_lib.TH_PAIRS]), n_entries=int(hdr[_lib.TH_ENTRIES]), sell_words=int(hdr[_lib.TH_SELL_WORDS]))
    o = a0 + 4 * _lib.TH_WORDS
    v["loci"] = blob[o:o + 4 * nl].view(np.uint32)
    o = a0 + _al16(o - a0 + 4 * nl)
    v["slots"] = blob[o:o + 4 * nl].view(np.uint32)
    o = a0 + _al16(o - a0 + 4 * nl)
    v["nplane"] = blob[o:o + 2 * npl].view(np.uint16)
    o = a0 + _al16(o - a0 + 2 * npl)
    v["count"] = blob[o:o + 8 * nc].view(np.float64)
    o = a0 + _al16(o - a0 + 8 * nc)
    n_words = int(((v["nplane"].astype(np.int64) + 3) // 4 * 4).sum())
    v["pairs"] = blob[o:o + 2 * n_words].view(np.uint16)
    assert _al16(o - a0 + 2 * n_words) == a_bytes
    b0 = a0 + a_bytes
    v["slices"] = blob[b0:b0 + 4 * ns].view(np.uint32)
    o = b0 + _al16(4 * ns)
    v["pos"] = blob[o:o + 2 * ni].view(np.uint16)
    o = b0 + _al16(o - b0 + 2 * ni)
    v["run_key"] = blob[o:o + 2 * nr].view(np.uint16)
    o = b0 + _al16(o - b0 + 2 * nr)
    v["run_first"] = blob[o:o + 2 * (nr + 1)].view(np.uint16)
    o = b0 + _al16(o - b0 + 2 * (nr + 1))
    v["ents"] = blob[o:o + 2 * v["sell_words"]].view(np.uint16)
    assert _al16(o - b0 + 2 * v["sell_words"]) == int(hdr[_lib.TH_B_BYTES])
    return v


def numerator_W(tiled, theta_T8, T, unit=False):
    """W[t][h] = sum over classes hitting (t, h) of count / normaliser, through the tile layout.  theta_T8: [T][8]."""
    blob, desc = tiled.arrays["blob"], tiled.arrays["tile_desc"].reshape(-1, _lib.TD_WORDS)
    partial = np.zeros((tiled.info["n_slots"], 8))
    written = np.zeros(tiled.info["n_slots"], dtype=bool)
    sub = np.zeros((T, 32))
    for m in range(16):
        for b in range(4):
            if (m >> b) & 1:
                sub[:, m] += 1.0 if unit else theta_T8[:, b]
                sub[:, 16 + m] += 1.0 if unit else theta_T8[:, 4 + b]
    seen_tiles = set()
    for row in desc[: tiled.info["n_tiles"]]:
        assert int(row[6]) not in seen_tiles
        seen_tiles.add(int(row[6]))
        v = tile_views(blob, row)
        nc, ni = v["n_classes"], v["n_items"]
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
        wts = np.append(v["count"] / s, 0.0)  # + the zero slot the padding entries point at
        assert sorted(v["pos"].tolist()) == list(range(ni)) and v["n_slices"] == (ni + 31) // 32
        isum = np.zeros(ni)
        real = 0
        last_len = 99
        for sidx, sw in enumerate(v["slices"]):
            first, L = int(sw) >> 5, int(sw) & 31
            assert L in (1, 2, 3, 4, 6, 8, 12, 16) and L <= last_len and first + 32 * L <= v["sell_words"]
            last_len = L
            block = v["ents"][first:first + 32 * L].reshape(L, 32)
            assert (block <= nc).all()
            real += int((block < nc).sum())
            sums = wts[block].sum(axis=0)
            for lane in range(32):
                vpos = sidx * 32 + lane
                if vpos < ni:
                    isum[v["pos"][vpos]] = sums[lane]
                else:
                    assert (block[:, lane] == nc).all()
        assert real == v["n_entries"]
        acc = np.zeros(v["n_loci"] * 32)
        rf = v["run_first"].astype(int)
        if ni:
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
