#!/usr/bin/env python3
"""Summarise a gpurun profiling pass (profiles/profile_*.sh) into text files that are committed under profiles/<tag>/.

  python profiles/summarize.py <tag> [launches_per_step]

Reads gpurun_out/launches.csv (ncu --metrics gpu__time_duration.sum: every launch of our kernels, cold-cache and serialised:
compare SHARES, not absolutes) and gpurun_out/prof_*.ncu-rep (ncu --set full captures, read with `ncu -i ... --page raw --csv`).
"""
import collections
import csv
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__cycles_elapsed.max", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed"]


def launches(tag, per_step, csv_name="launches.csv", out_name="launches_step.txt"):
    path = os.path.join(ROOT, "gpurun_out", csv_name)
    if not os.path.exists(path):
        return
    rows, hdr = [], None
    with open(path) as f:
        for ln in f:
            if ln.startswith('"ID"'):
                hdr = next(csv.reader([ln]))
                break
        for r in csv.reader(f):
            if hdr and len(r) == len(hdr):
                rows.append(dict(zip(hdr, r)))
    # a step is one period of the launch sequence: the distance between the last two MAC launches
    pre = [r for r in rows[:len(rows) - 1] if ("k_encode" in r["Kernel Name"] or "k_geno" in r["Kernel Name"])]
    if "otf" in csv_name:
        pre = []
    # a step starts with the rotation-by-0 copies of the baby phase: the k_copy_add launch that follows the previous step's last k_md_final
    starts = [i for i, r in enumerate(rows) if "k_copy_add" in r["Kernel Name"] and (i == 0 or "k_md_final" in rows[i - 1]["Kernel Name"] or "k_img" in rows[i - 1]["Kernel Name"] or "k_encode" in rows[i - 1]["Kernel Name"])]
    macs = [i for i, r in enumerate(rows) if "k_mac" in r["Kernel Name"]]
    if per_step <= 0:
        per_step = len(rows) - starts[-1] if starts else (macs[-1] - macs[-2] if len(macs) >= 2 else len(rows))
    nsteps = len(starts) if starts else len(macs)
    step = rows[len(rows) - per_step:]
    agg = collections.OrderedDict()
    for r in step:
        nm = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "")
        a = agg.setdefault((nm, r["Grid Size"], r["Block Size"]), [0, 0.0])
        a[0] += 1
        a[1] += float(r["Metric Value"]) / 1e6
    tot = sum(a[1] for a in agg.values())
    out = ["# launch list of ONE step (the last of %d captured; %d launches/step), ncu gpu__time_duration.sum, --clock-control none" % (nsteps, per_step),
           "# cold-cache + serialised per-launch times: compare shares, not absolutes.  total %.2f ms" % tot,
           "%-44s %-16s %-12s %5s %10s %7s %10s" % ("kernel", "grid", "block", "n", "ms", "share", "avg_us")]
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append("%-44s %-16s %-12s %5d %10.3f %6.1f%% %10.1f" % (k[0][:44], k[1], k[2], a[0], a[1], 100 * a[1] / tot, 1e3 * a[1] / a[0]))
    if pre:
        out.append("# preprocess launches (outside the step):")
        for r in pre[:40]:
            out.append("%-44s %-16s %10.3f ms" % (re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "")[:44], r["Grid Size"], float(r["Metric Value"]) / 1e6))
    d = os.path.join(ROOT, "profiles", tag)
    os.makedirs(d, exist_ok=True)
    with open(os.path.join(d, out_name), "w") as f:
        f.write("\n".join(out) + "\n")
    print("\n".join(out))


def reports(tag):
    d = os.path.join(ROOT, "profiles", tag)
    os.makedirs(d, exist_ok=True)
    for rep in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", "prof_*.raw.csv"))):
        name = os.path.basename(rep)[5:-8]
        r = list(csv.reader(open(rep).read().splitlines()))
        if len(r) < 3:
            continue
        h, units = r[0], r[1]
        out = ["# ncu --set full --clock-control none, kernel regex %s; source: gpurun_out/%s" % (name, os.path.basename(rep))]
        for row in r[2:]:
            out.append("kernel: " + row[h.index("Kernel Name")][:160])
            out.append("  grid %s block %s" % (row[h.index("Grid Size")], row[h.index("Block Size")]))
            for k in KEYS:
                if k in h:
                    out.append("  %-80s %s %s" % (k, row[h.index(k)], units[h.index(k)]))
        with open(os.path.join(d, "ncu_%s.txt" % name), "w") as f:
            f.write("\n".join(out) + "\n")
        print("\n".join(out))


if __name__ == "__main__":
    tag = sys.argv[1]
    per_step = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    launches(tag, per_step)
    launches(tag, 0, "launches_T.csv", "launches_step_transposed.txt")
    launches(tag, 0, "launches_otf13.csv", "launches_step_on_the_fly_logN13.txt")
    reports(tag)
