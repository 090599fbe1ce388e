#!/usr/bin/env python
"""Turn the scratch outputs of a GPU session (gpurun_out/) into the committed evidence under profiles/.

    python tools/make_profiles.py r01
"""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
os.makedirs(PROF, exist_ok=True)


def short(name):
    return name.replace("void ", "").replace("<unnamed>::", "").split("(")[0]


def launches():
    rows = list(csv.reader(open(os.path.join(OUT, "launches.csv"))))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    c = {h: i for i, h in enumerate(hdr)}
    per_id = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) < len(hdr):
            continue
        d = per_id.setdefault(r[c["ID"]], {"kernel": short(r[c["Kernel Name"]]), "grid": r[c["Grid Size"]], "block": r[c["Block Size"]]})
        v = float(r[c["Metric Value"]].replace(",", ""))
        unit = r[c["Metric Unit"]]
        name = r[c["Metric Name"]]
        if name == "gpu__time_duration.sum":
            d["ms"] = v / {"ns": 1e6, "us": 1e3, "ms": 1.0, "s": 1e-3}.get(unit, 1e6)
        else:
            mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1.0)
            d[name] = v * mult
    return list(per_id.values())


L = launches()
with open(os.path.join(PROF, f"{tag}_launches.csv"), "w", newline="") as fh:
    w = csv.writer(fh)
    w.writerow(["#", "kernel", "grid", "block", "gpu_time_ms", "dram_read_bytes", "dram_write_bytes"])
    for i, d in enumerate(L):
        w.writerow([i, d["kernel"], d["grid"], d["block"], f"{d.get('ms', 0):.4f}", int(d.get("dram__bytes_read.sum", 0)),
                    int(d.get("dram__bytes_write.sum", 0))])
agg = collections.OrderedDict()
for d in L:
    a = agg.setdefault(d["kernel"], {"n": 0, "ms": 0.0, "rd": 0.0, "wr": 0.0})
    a["n"] += 1
    a["ms"] += d.get("ms", 0)
    a["rd"] += d.get("dram__bytes_read.sum", 0)
    a["wr"] += d.get("dram__bytes_write.sum", 0)
tot = sum(a["ms"] for a in agg.values())
lines = [f"# {tag}: ncu launch list of `python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu` (2 steps + input generation)",
         "", "`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none` "
         "(cold-cache, serialised: compare SHARES, not absolutes).  Full list: `%s_launches.csv`." % tag, "",
         "| kernel | launches | total ms | share | DRAM read GB | DRAM write GB |", "|---|---:|---:|---:|---:|---:|"]
for k, a in agg.items():
    lines.append(f"| `{k}` | {a['n']} | {a['ms']:.3f} | {a['ms'] / tot * 100:.1f}% | {a['rd'] / 1e9:.2f} | {a['wr'] / 1e9:.2f} |")
setop = [d for d in L if d["kernel"].startswith(("setop_", "search_partition", "nway_"))]
so_ms = sum(d.get("ms", 0) for d in setop)
lines += ["", f"set-operation kernels (N-way union + its partition, two-way passes, searches, partitions): {so_ms:.2f} ms of {tot:.2f} ms = "
          f"{so_ms / tot * 100:.1f}% of all GPU time in the run (the rest is the synthetic input generator `select_kernel<MemberGen>`, "
          "outside the timed region)."]
open(os.path.join(PROF, f"{tag}_launch_summary.md"), "w").write("\n".join(lines) + "\n")


# dram traffic per stats scope of bench.py (= one N-way operation, or one two-way pass of inter / diff)
def dram(ds):
    return sum(d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0) for d in ds)


def fam(pred):
    return [d for d in L if pred(d["kernel"])]


traffic = {}
nw_union = fam(lambda k: k.startswith("nway_kernel<0"))
if nw_union:
    # the partition / check launches belong to the union calls (inter / diff run file by file by default)
    aux = fam(lambda k: k.startswith(("nway_partition_kernel", "nway_check_kernel")))
    traffic["setop_union_nway"] = {"dram_bytes_per_launch": (dram(nw_union) + dram(aux)) / len(nw_union), "launches": len(nw_union),
                                   "kernels": "nway_partition_kernel x3 + nway_check_kernel + nway_kernel<UNION>"}
for op, name in ((0, "setop_inter"), (1, "setop_diff"), (2, "setop_union")):
    main = fam(lambda k, op=op: k.startswith((f"setop_pipe_kernel<{op},", f"setop_search_kernel<{op},", f"setop_fast_kernel<{op},", f"setop_kernel<{op},")))
    if main:
        traffic[name] = {"dram_bytes_per_launch": dram(main) / len(main), "launches": len(main),
                         "kernels": "two-way pass kernels only (their partition kernels are shared between inter and diff in the list)"}
traffic["source"] = f"profiles/{tag}_launches.csv (dram__bytes_read.sum + dram__bytes_write.sum per launch, summed per operation / pass)"
json.dump(traffic, open(os.path.join(PROF, "setop_ncu_traffic.json"), "w"), indent=1)

# --set full captures -> key metrics
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
KEYS += ["l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"]
for rep, title in (("nway_union_prof", "nway_kernel<UNION> (one 8-way union of the C3 files, 4e9 k-mers in)"),
                   ("nway_filter_prof", "nway_kernel<INTER> (opt-in N-way hash filter, 4e9 k-mers in)"),
                   ("setop_pipe_prof", "setop_pipe_kernel (first two-way passes of inter / diff)"),
                   ("setop_search_prof", "setop_search_kernel (a later pass of inter: running set looked up in the next file)"),
                   ("setop_union_prof", "setop_pipe_kernel (union passes, 1e9 k-mers in)"), ("onesweep_prof", "onesweep_kernel (one 8-bit pass over 3e8 keys)"),
                   ("kmer_prof", "kmer_kernel<HASHED> (ntHash k=31 canonical over 3e8 bases: iterator, then the filtered variant of count)"),
                   ("fold_prof", "fold_kernel<UNIQUE> (2.5e8 sorted keys)"), ("minimizer_prof", "minimizer_kernel (w=15 over 3e8 hashes)")):
    path = os.path.join(OUT, rep + ".ncu-rep")
    if not os.path.exists(path):
        continue
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    out = [f"# {tag}: `ncu --set full --clock-control none --import-source on` -- {title}", ""]
    out += ["| metric | unit | " + " | ".join(f"launch {i}" for i in range(len(rows) - 2)) + " |", "|---|---|" + "---:|" * (len(rows) - 2)]
    for k in ["Kernel Name"] + KEYS:
        if k in hdr:
            i = hdr.index(k)
            out.append(f"| `{k}` | {units[i]} | " + " | ".join(short(r[i])[:60] if k == "Kernel Name" else r[i] for r in rows[2:]) + " |")
    sass = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass", "--kernel-id", ":::1"],
                          capture_output=True, text=True).stdout
    tmp = os.path.join(OUT, rep + "_sass.csv")
    open(tmp, "w").write(sass)
    summ = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_sass_summary.py"), tmp, "12"], capture_output=True, text=True).stdout
    out += ["", "SASS-level summary of launch 1 (tools/ncu_sass_summary.py; the listing holds the kernel twice, so counts are doubled):", "", "```", summ.strip(), "```"]
    mn = subprocess.run(f"cuobjdump -sass {os.path.join(ROOT, 'unikmer_b200', 'libukm.so')} | grep -oE 'UBLKCP[.A-Z0-9]*|SYNCS[.A-Z0-9]*|LDG.E.128[.A-Z]*|STG.E.128|MATCH.ANY|VOTE[.A-Z]*|POPC' | sort | uniq -c",
                        shell=True, capture_output=True, text=True).stdout
    out += ["", "Blackwell-native mnemonics in libukm.so (`cuobjdump -sass`): TMA bulk copies = `UBLKCP`, mbarrier = `SYNCS.*`:", "", "```", mn.strip(), "```"]
    open(os.path.join(PROF, f"{tag}_{rep}.md"), "w").write("\n".join(out) + "\n")

for src, dst in (("bench_full.json", f"{tag}_bench.json"), ("bench_ref.json", f"{tag}_bench_reference.json"), ("microbench.jsonl", f"{tag}_microbench.jsonl"),
                 ("pytest_gpu.log", f"{tag}_pytest_gpu.log"), ("exp_nway.jsonl", f"{tag}_exp_nway.jsonl")):
    if os.path.exists(os.path.join(OUT, src)):
        shutil.copy(os.path.join(OUT, src), os.path.join(PROF, dst))
print("profiles written:", sorted(os.listdir(PROF)))
