#!/usr/bin/env python
"""Splits the warp samples of an .ncu-rep (first kernel, SASS view) into code regions near HMMA instructions (tensor-core
tile loops) and the rest, and lists the hottest instructions of the large 'other' regions.
usage: python tools/ncu_regions.py report.ncu-rep [window=60]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; W = int(sys.argv[2]) if len(sys.argv) > 2 else 60
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src))); h = rows[1]; data = rows[2:]
isrc = h.index('Source'); isamp = h.index('# Samples'); iinst = h.index('Instructions Executed')
stalls = [i for i, c in enumerate(h) if c.startswith('stall_') and 'Not Issued' not in c]
n = len(data)
is_h = ['HMMA' in r[isrc] for r in data]
near = [False] * n
last = -10 ** 9
for i in range(n):
    if is_h[i]: last = i
    if i - last <= W: near[i] = True
nxt = 10 ** 9
for i in range(n - 1, -1, -1):
    if is_h[i]: nxt = i
    if nxt - i <= W: near[i] = True
S = lambda r: int(r[isamp] or 0)
tot = sum(S(r) for r in data)
mm = sum(S(r) for r, f in zip(data, near) if f)
print(f'samples: total {tot}, near HMMA {mm} ({100 * mm / tot:.1f} %), other {tot - mm} ({100 * (tot - mm) / tot:.1f} %)')
segs = []; cur = near[0]; s = 0; start = 0
for i, (r, f) in enumerate(zip(data, near)):
    if f != cur:
        segs.append((cur, start, i, s)); cur = f; s = 0; start = i
    s += S(r)
segs.append((cur, start, n, s))
for f, a, b, s in segs:
    if s > 0.01 * tot:
        print(f"{'MMA  ' if f else 'other'} SASS[{a}:{b}] samples {s} ({100 * s / tot:.1f} %)")
        if not f:
            for k, r in sorted(enumerate(data[a:b]), key=lambda x: -S(x[1]))[:4]:
                st = sorted([(int(r[i] or 0), h[i]) for i in stalls], reverse=True)[:1]
                print(f'        {S(r):6d} x{r[iinst]:>8s}  {r[isrc][:56]:56s} {st}')
