#!/usr/bin/env python
"""Group the SASS page of an ncu report (`ncu -i x.ncu-rep --page source --csv --print-source sass`) into runs of
instructions with the same execution count (= loop bodies / straight-line regions) and print their share of all executed
warp instructions, plus the stall mix."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
minshare = float(sys.argv[2]) if len(sys.argv) > 2 else 0.006
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
hdr = rows[hi]
col = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
ex = [float(r[col['Instructions Executed']] or 0) for r in data]
sm = [float(r[col['# Samples']] or 0) for r in data]
tot = sum(ex)
print('total warp-instr %.3e  samples %d' % (tot, sum(sm)))
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
mix = {s: sum(float(r[col[s]] or 0) for r in data) for s in stalls}
print('stalls:', ', '.join(f'{k[6:]}={v / max(sum(sm), 1) * 100:.1f}%' for k, v in sorted(mix.items(), key=lambda kv: -kv[1]) if v > 0.005 * sum(sm)))
print('shared wavefronts: total %.3e ideal %.3e' % (sum(float(r[col['L1 Wavefronts Shared']] or 0) for r in data),
                                                    sum(float(r[col['L1 Wavefronts Shared Ideal']] or 0) for r in data)))
regs, start = [], 0
for i in range(1, len(data) + 1):
    if i == len(data) or abs(ex[i] - ex[start]) > 0.02 * max(ex[start], 1):
        regs.append((start, i - 1, ex[start], sum(ex[start:i]), sum(sm[start:i])))
        start = i
cov = 0
for s, e, c, t, smp in regs:
    if t > minshare * tot:
        cov += t
        thr = float(data[s][col['Avg. Threads Executed']] or 0)
        print(f'[{s:4d}-{e:4d}] n={e - s + 1:3d} exec={c:12.0f} instr%={t / tot * 100:5.1f} samples%={smp / max(sum(sm), 1) * 100:5.1f} thr={thr:4.1f}  {data[s][col["Source"]].strip()[:56]}')
print('covered %.2f' % (cov / tot))
