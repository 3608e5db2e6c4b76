"""Host-side model of the gather locality of a class order (no GPU needed): for the C2 benchmark shape it counts what the
two passes of an EM update would request under a given order of the alignment classes --

  column pass  distinct 32-byte sectors / 128-byte lines of the weight vector gathered per locus (entries of a locus are
               visited in class order; deep loci in two parts, partial and full masks)
  row pass     distinct subset-table rows (loci) per warp-wide table load (32 consecutive classes of one width, one pair slot)

Orders: `current` = (width, smallest locus), `second` = + second-smallest locus (the shipped order since round 2),
`third` = + third-smallest, `window` = classes grouped into windows of ~`--window` consecutive classes by smallest locus
first and by width inside a window (the proposal of DESIGN.md section 10, item 0: needs per-(window, width) segment
descriptors in the row pass).      python tools/locality_model.py [--classes 5000000]
Results of the run recorded in profiles/r2_locality_model.txt."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gbrs_b200 import synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--loci", type=int, default=80000)
    ap.add_argument("--classes", type=int, default=5_000_000)
    ap.add_argument("--window", type=int, default=2048)
    a = ap.parse_args()
    d = synth.generate(T=a.loci, N=a.classes, H=8)
    pc, pl, pm, N = d.pair_class, d.pair_locus, d.pair_mask, d.N
    k = np.bincount(pc, minlength=N)
    start = np.cumsum(k) - k
    last = len(pl) - 1
    minloc = pl[start]
    sec = np.where(k > 1, pl[np.minimum(start + 1, last)], 0)
    third = np.where(k > 2, pl[np.minimum(start + 2, last)], 0)
    width = np.minimum(k, 9)
    c = np.arange(N)
    # windows: classes sorted by smallest locus, cut every `window` classes (a window never splits a locus' classes unevenly
    # enough to matter for this count)
    by_loc = np.argsort(minloc, kind="stable")
    win = np.empty(N, np.int64)
    win[by_loc] = np.arange(N) // a.window
    cnt = np.bincount(pl, minlength=d.T)
    deep = cnt[pl] > 512
    part = np.where(deep & (pm == 255), 1, 0)
    print(f"T={d.T} N={N} pairs={d.pairs}; lower bounds: {d.pairs // 4} sectors, {d.pairs // 16} lines")

    def analyse(keys, name, segmented=False):
        order = np.lexsort(keys)  # last key is the primary one
        new_id = np.empty(N, np.int64)
        new_id[order] = np.arange(N)
        nid = new_id[pc]
        o = np.lexsort((nid, part, pl))
        seg = pl[o] * 2 + part[o]
        ids = nid[o]
        newseg = np.r_[True, seg[1:] != seg[:-1]]
        sectors = (np.r_[True, ids[1:] // 4 != ids[:-1] // 4] | newseg).sum()
        lines = (np.r_[True, ids[1:] // 16 != ids[:-1] // 16] | newseg).sum()
        # row pass: warps of 32 consecutive classes of one width (per window when segmented), distinct loci per pair slot
        tot = reqs = 0
        grp = (win[order] * 16 + width[order]) if segmented else width[order]
        bounds = np.flatnonzero(np.r_[True, grp[1:] != grp[:-1], True])
        warp_id = np.zeros(N, np.int64)
        warp_id[bounds[:-1]] = 1  # a new group starts a new warp
        pos_in_grp = np.arange(N) - np.repeat(bounds[:-1], np.diff(bounds))
        warp_id = np.cumsum(warp_id | (pos_in_grp % 32 == 0))
        lanes_used = N / (warp_id[-1] * 32)
        for p in range(8):
            has = k[order] > p
            w = warp_id[has]
            loc = pl[start[order][has] + p]
            o2 = np.lexsort((loc, w))
            w2, l2 = w[o2], loc[o2]
            tot += (np.r_[True, (w2[1:] != w2[:-1]) | (l2[1:] != l2[:-1])]).sum()
            reqs += len(np.unique(w))
        print(f"{name:8s} column: {sectors / 1e6:6.2f} M sectors {lines / 1e6:5.2f} M lines | row: {tot / reqs:5.2f} distinct table rows per "
              f"warp load, {reqs / 1e3:6.0f} k warp loads, lane use {lanes_used:.2f}")

    analyse((c, minloc, width), "current")
    analyse((c, sec, minloc, width), "second")
    analyse((c, third, sec, minloc, width), "third")
    analyse((c, sec, minloc, width, win), "window", segmented=True)
    for wsize in (512, 128):
        win[by_loc] = np.arange(N) // wsize
        analyse((c, sec, minloc, width, win), f"win{wsize}", segmented=True)
    win[by_loc] = np.arange(N)  # every class its own window: pure locus order, widths mixed (a row pass with row pointers)
    analyse((c, third, sec, minloc), "locus", segmented=True)


if __name__ == "__main__":
    main()
