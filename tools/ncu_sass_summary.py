#!/usr/bin/env python
"""Summarise `ncu --page source --csv --print-source sass` output: stall mix, hottest instructions."""
import csv
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hi]
data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
col = {h: i for i, h in enumerate(hdr)}


def f(r, name):
    try:
        return float(r[col[name]].replace(",", "") or 0)
    except ValueError:
        return 0.0


tot_samples = sum(f(r, "# Samples") for r in data)
tot_inst = sum(f(r, "Instructions Executed") for r in data)
print(f"SASS instructions: {len(data)}  warp-inst executed: {tot_inst:.3e}  samples: {tot_samples:.0f}")
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
mix = {s: sum(f(r, s) for r in data) for s in stalls}
print("stall mix (all samples):", ", ".join(f"{k[6:]}={v / max(tot_samples, 1) * 100:.1f}%" for k, v in sorted(mix.items(), key=lambda kv: -kv[1]) if v))
print("shared wavefronts: total", sum(f(r, "L1 Wavefronts Shared") for r in data), "ideal", sum(f(r, "L1 Wavefronts Shared Ideal") for r in data))
print(f"\n{'idx':>5} {'samples%':>8} {'exec':>10} {'thr':>5}  sass")
order = sorted(range(len(data)), key=lambda i: -f(data[i], "# Samples"))[:top]
for i in sorted(order):
    r = data[i]
    print(f"{i:5d} {f(r, '# Samples') / max(tot_samples, 1) * 100:8.2f} {f(r, 'Instructions Executed'):10.0f} {f(r, 'Avg. Threads Executed'):5.1f}  {r[col['Source']][:110]}")
# executed-instruction histogram by opcode
ops = {}
for r in data:
    op = r[col["Source"]].split()[0] if r[col["Source"]].split() else "?"
    if op.startswith("@"):
        op = r[col["Source"]].split()[1]
    op = op.split(".")[0]
    ops[op] = ops.get(op, 0) + f(r, "Instructions Executed")
print("\nexecuted by opcode:", ", ".join(f"{k}={v / tot_inst * 100:.1f}%" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:18]))
