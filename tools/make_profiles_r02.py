#!/usr/bin/env python
"""Turn the scratch outputs of the round-2 GPU sessions (gpurun_out/, written by tools/gpu_evidence_r02.sh and the
multi-GPU runs) into the committed evidence under profiles/ (r02_*).

    python tools/make_profiles_r02.py
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
tag = "r02"


def short(name):
    return name.replace("void ", "").replace("<unnamed>::", "").split("(")[0]


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    c = {h: i for i, h in enumerate(hdr)}
    per_id = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) < len(hdr):
            continue
        d = per_id.setdefault(r[c["ID"]], {"kernel": short(r[c["Kernel Name"]]), "grid": r[c["Grid Size"]], "block": r[c["Block Size"]]})
        v = float(r[c["Metric Value"]].replace(",", ""))
        unit, name = r[c["Metric Unit"]], r[c["Metric Name"]]
        if name == "gpu__time_duration.sum":
            d["ms"] = v / {"ns": 1e6, "us": 1e3, "ms": 1.0, "s": 1e-3}.get(unit, 1e6)
        else:
            d[name] = v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1.0)
    return list(per_id.values())


lp = os.path.join(OUT, "launches.csv")
if os.path.exists(lp):
    L = launches(lp)
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
    lines = [f"# {tag}: ncu launch list of `python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu`",
             "", "`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none` (cold-cache, serialised: "
             f"compare SHARES, not absolutes).  The run holds the warm-up step, the timed step, the three-separate-calls comparison (4 x 3 calls) and "
             f"the input generator.  Full list: `{tag}_launches.csv`.", "",
             "| kernel | launches | total ms | share | DRAM read GB | DRAM write GB |", "|---|---:|---:|---:|---:|---:|"]
    for k, a in agg.items():
        lines.append(f"| `{k}` | {a['n']} | {a['ms']:.3f} | {a['ms'] / tot * 100:.1f}% | {a['rd'] / 1e9:.2f} | {a['wr'] / 1e9:.2f} |")

    def dram(ds):
        return sum(d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0) for d in ds)
    # the fused call = nway partition x3 + check + nway_kernel<0,8> (with the riding inter / diff) + 2 x (count, scan, gather)
    traffic = {}
    nwk = [d for d in L if d["kernel"].startswith("nway_kernel<0")]
    # the launches with inter / diff riding along are the ones followed by nfilter_count_kernel without an nfilter_kernel in between
    idx = {id(d): i for i, d in enumerate(L)}
    fused, plain = [], []
    for d in nwk:
        i = idx[id(d)]
        nxt = L[i + 1]["kernel"] if i + 1 < len(L) else ""
        (fused if nxt.startswith("nfilter_count_kernel") else plain).append(d)
    part = [d for d in L if d["kernel"].startswith(("nway_partition_kernel", "nway_check_kernel"))]
    n_union_calls = max(len(nwk), 1)
    if fused:
        traffic["setop_inter_diff_union_nway"] = {
            "dram_bytes_per_launch": dram(fused) / len(fused) + dram(part) / n_union_calls, "launches": len(fused),
            "kernel_only_dram_bytes_per_launch": dram(fused) / len(fused),
            "kernels": "nway_partition_kernel x3 + nway_check_kernel + nway_kernel<UNION> with inter / diff riding along"}
    if plain:
        traffic["setop_union_nway"] = {"dram_bytes_per_launch": dram(plain) / len(plain) + dram(part) / n_union_calls, "launches": len(plain),
                                       "kernel_only_dram_bytes_per_launch": dram(plain) / len(plain),
                                       "kernels": "nway_partition_kernel x3 + nway_check_kernel + nway_kernel<UNION>"}
    for op, name in ((0, "setop_inter_nway"), (1, "setop_diff_nway"), (2, "setop_inter_diff_nway")):
        main = [d for d in L if d["kernel"].startswith(f"nfilter_kernel<{op},")]
        if main:
            npart = [d for d in L if d["kernel"].startswith("nfilter_partition_kernel")]
            all_nf = [d for d in L if d["kernel"].startswith("nfilter_kernel<")]
            traffic[name] = {"dram_bytes_per_launch": dram(main) / len(main) + dram(npart) / max(len(all_nf), 1), "launches": len(main),
                             "kernel_only_dram_bytes_per_launch": dram(main) / len(main),
                             "kernels": "nfilter_partition_kernel x2 + nfilter_kernel (the mask count / scan / gather passes add ~0.6 GB)"}
    traffic["source"] = f"profiles/{tag}_launches.csv (dram__bytes_read.sum + dram__bytes_write.sum per launch, summed per operation)"
    json.dump(traffic, open(os.path.join(PROF, "setop_ncu_traffic.json"), "w"), indent=1)
    lines += ["", "DRAM traffic per operation (profiles/setop_ncu_traffic.json, read by bench.py as `roofline.traffic`):", ""]
    for k, v in traffic.items():
        if isinstance(v, dict):
            lines.append(f"* `{k}`: {v['dram_bytes_per_launch'] / 1e9:.2f} GB per call ({v['kernels']})")
    # one fused step, launch by launch: the share of every kernel in the step (to be compared with bench.py's
    # roofline.share_of_step, which comes from CUDA events in an unprofiled run)
    if fused:
        i1 = idx[id(fused[-1])]
        i0 = i1
        while i0 > 0 and L[i0 - 1]["kernel"].startswith(("nway_partition_kernel", "nway_check_kernel")):
            i0 -= 1
        i2 = i1
        while i2 + 1 < len(L) and L[i2 + 1]["kernel"].startswith(("nfilter_count_kernel", "nfilter_scan_kernel", "nfilter_gather_kernel")):
            i2 += 1
        step = L[i0:i2 + 1]
        st = sum(d.get("ms", 0) for d in step)
        lines += ["", f"One step of the bench (= one `ukm_setops_stream` call: launches {i0}..{i2} of the list), under ncu:", "",
                  "| # | kernel | ms | share of the step | DRAM read GB | DRAM write GB |", "|---:|---|---:|---:|---:|---:|"]
        for j, d in enumerate(step):
            lines.append(f"| {i0 + j} | `{d['kernel']}` | {d.get('ms', 0):.3f} | {d.get('ms', 0) / st * 100:.1f}% | "
                         f"{d.get('dram__bytes_read.sum', 0) / 1e9:.2f} | {d.get('dram__bytes_write.sum', 0) / 1e9:.2f} |")
        fam = sum(d.get("ms", 0) for d in step if d["kernel"].startswith(("nway_", )))
        lines += ["", f"Sum {st:.3f} ms; the fused kernel with its partition (the family bench.py reports as `setop_inter_diff_union_nway`): "
                  f"{fam:.3f} ms = {fam / st * 100:.1f}% of the step's kernel time (bench.py, CUDA events, unprofiled: share_of_step 0.979 of the "
                  "step's wall time including the host-side gaps)."]
    open(os.path.join(PROF, f"{tag}_launch_summary.md"), "w").write("\n".join(lines) + "\n")

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
for rep, title in (("nway_union3_prof", "nway_kernel<UNION> with inter / diff riding along (ONE pass for the three results of C3, 4e9 k-mers in)"),
                   ("nway_union_prof", "nway_kernel<UNION>, plain (one 8-way union of the C3 files)"),
                   ("nfilter_prof", "nfilter_kernel<INTER> (single-pass 8-way inter over file-0 chunks, C3)"),
                   ("r2_nunion_prof2", "nunion_kernel<8> (row-based union, opt-in: the experiment of DESIGN.md 4.1b)")):
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
    sass = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    tmp = os.path.join(OUT, rep + "_sass.csv")
    open(tmp, "w").write(sass)
    summ = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_regions.py"), tmp, "0.02"], capture_output=True, text=True).stdout
    out += ["", "SASS regions (tools/ncu_regions.py: runs of instructions with one execution count = loop bodies; share of executed warp instructions):",
            "", "```", summ.strip(), "```"]
    open(os.path.join(PROF, f"{tag}_{rep.replace('r2_', '')}.md"), "w").write("\n".join(out) + "\n")
mn = subprocess.run(f"cuobjdump -sass {os.path.join(ROOT, 'unikmer_b200', 'libukm.so')} | grep -oE 'UBLKCP[.A-Z0-9]*|SYNCS[.A-Z0-9]*|REDUX[.A-Z0-9]*|VIADDMNMX[.A-Z0-9]*|ATOMS[.A-Z0-9]*' | sort | uniq -c",
                    shell=True, capture_output=True, text=True).stdout
open(os.path.join(PROF, f"{tag}_sass_mnemonics.txt"), "w").write(
    "Blackwell-native mnemonics in libukm.so (cuobjdump -sass): TMA bulk copies = UBLKCP, mbarrier = SYNCS.*, warp reductions = REDUX\n\n" + mn)

for src, dst in (("bench_full.json", f"{tag}_bench.json"), ("bench_ref.json", f"{tag}_bench_reference.json"), ("microbench.jsonl", f"{tag}_microbench.jsonl"),
                 ("pytest_gpu.log", f"{tag}_pytest_gpu.log"), ("exp_ops.jsonl", f"{tag}_exp_ops.jsonl"), ("c4.json", f"{tag}_c4.json"),
                 ("c5_n8.json", f"{tag}_c5.json"), ("gpu_info.csv", f"{tag}_gpu_info.csv"),
                 ("bench_n2.json", f"{tag}_bench_n2.json"), ("bench_n4.json", f"{tag}_bench_n4.json"), ("bench_n8.json", f"{tag}_bench_n8.json")):
    if os.path.exists(os.path.join(OUT, src)):
        shutil.copy(os.path.join(OUT, src), os.path.join(PROF, dst))
print("profiles written:", sorted(f for f in os.listdir(PROF) if f.startswith(tag)))
