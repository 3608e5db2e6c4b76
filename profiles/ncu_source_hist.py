"""Opcode histogram (weighted by executed count) and stall hot spots from `ncu --page source --csv`.
usage: python profiles/ncu_source_hist.py src.csv"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
isrc, ist, iex = hdr.index('Source'), hdr.index('Warp Stall Sampling (All Samples)'), hdr.index('Instructions Executed')
seen, data = set(), []
for r in rows[2:]:
    if len(r) > iex and r[iex].isdigit() and r[0] not in seen:
        seen.add(r[0])
        data.append((int(r[iex]), int(r[ist]), r[isrc].strip()))
tot = sum(d[0] for d in data)
tots = max(sum(d[1] for d in data), 1)
print(rows[0][1][:80], '| warp instructions', tot, '| samples', tots, '| SASS lines', len(data))
h, hs = collections.Counter(), collections.Counter()
for ex, st, src in data:
    parts = src.split()
    op = (parts[1] if src.startswith('@') else parts[0]).split('.')[0]
    h[op] += ex
    hs[op] += st
for op, c in h.most_common(18):
    print(f"  {op:10s} {c:10d} {100 * c / tot:5.1f}%   stall {100 * hs[op] / tots:5.1f}%")
print('  --- top stall lines')
for ex, st, src in sorted(data, key=lambda d: -d[1])[:10]:
    print(f"  {st:6d} {ex:9d}  {src[:100]}")
