#!/usr/bin/env python
"""Summarise one gpurun profiling pass into profiles/ (tracked).

    python scripts/summarize_profile.py <tag> [<round-name>]

Reads gpurun_out/launches_<tag>.csv (ncu --metrics gpu__time_duration.sum launch
list), gpurun_out/prof_<tag>.ncu-rep (ncu --set full of the apply kernel) and
gpurun_out/bench_<tag>.log; writes profiles/<round>_launches.csv,
profiles/<round>_apply_full.txt (key counters, stall mix, instruction mix) and
updates profiles/traffic.json (dram bytes per launch, read by bench.py)."""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
rnd = sys.argv[2] if len(sys.argv) > 2 else tag
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed.avg.per_cycle_elapsed", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "launch__block_size", "launch__grid_size", "launch__shared_mem_per_block_dynamic",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.max"]

lst = os.path.join(G, f"launches_{tag}.csv")
if os.path.exists(lst):
    rows = [r for r in csv.reader(open(lst)) if len(r) > 10]
    with open(os.path.join(P, f"{rnd}_launches.csv"), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none of: python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e\n")
        f.write("id,kernel,block,grid,ns\n")
        tot = collections.Counter()
        for r in rows[1:]:
            f.write(f"{r[0]},\"{r[4]}\",\"{r[7]}\",\"{r[8]}\",{r[-1]}\n")
            tot[r[4].split('(')[0]] += float(r[-1])
        s = sum(tot.values())
        for k, v in tot.most_common():
            f.write(f"# share {k}: {v / s * 100:.1f}% ({v / 1e6:.3f} ms total)\n")

rep = os.path.join(G, f"prof_{tag}.ncu-rep")
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (vals[i], units[i]) for i, h in enumerate(hdr)}
    out = [f"# ncu --set full --clock-control none --import-source on, one launch of {d.get('Kernel Name', ('?',))[0]}"]
    for k in KEYS:
        if k in d:
            out.append(f"{k:70s} {d[k][0]:>18s} {d[k][1]}")
    def num(k):
        v, u = d[k]
        v = float(v.replace(",", ""))
        return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}.get(u, 1)
    traffic = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
    out.append(f"dram traffic per launch (read+write) = {traffic / 1e9:.3f} GB")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    h2 = rows[1]
    ix = {h: i for i, h in enumerate(h2)}
    cols = [h for h in h2 if h.startswith("stall_") and "Not Issued" not in h]
    st, ops, total = collections.Counter(), collections.Counter(), 0
    for r in rows[2:]:
        if len(r) < len(h2):
            continue
        for c in cols:
            st[c] += int(r[ix[c]])
        t = r[ix["Source"]].split()
        op = t[1] if t[0].startswith("@") else t[0]
        op = ".".join(op.split(".")[:2]) if op[:2] in ("LD", "ST", "UB") else op.split(".")[0]
        n = int(r[ix["Instructions Executed"]])
        ops[op] += n
        total += n
    out.append("\n# warp stall sampling (all samples)")
    s = sum(st.values()) or 1
    for c, n in st.most_common(8):
        out.append(f"{c:28s} {n / s * 100:6.2f}%")
    out.append(f"\n# instruction mix (warp instructions, total {total})")
    for c, n in ops.most_common(16):
        out.append(f"{c:14s} {n / total * 100:6.2f}%")
    open(os.path.join(P, f"{rnd}_apply_full.txt"), "w").write("\n".join(out) + "\n")
    tj = os.path.join(P, "traffic.json")
    t = json.load(open(tj)) if os.path.exists(tj) else {}
    t["apply_L32_f64_bytes_per_launch"] = traffic
    t["source"] = f"profiles/{rnd}_apply_full.txt"
    json.dump(t, open(tj, "w"), indent=1)
    print("\n".join(out))

b = os.path.join(G, f"bench_{tag}.log")
if os.path.exists(b):
    shutil.copy(b, os.path.join(P, f"{rnd}_bench.json"))
