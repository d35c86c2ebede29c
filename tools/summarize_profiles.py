#!/usr/bin/env python3
"""Turns raw gpurun_out/ captures into the tracked summaries under profiles/.

  python tools/summarize_profiles.py launches <launches.csv> <out.md> [title]
  python tools/summarize_profiles.py full <report.ncu-rep> <out.txt> <traffic-key>
"""
import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def short_name(name):
    m = re.search(r"ps_kernel<(?:\(anonymous namespace\)::|<unnamed>::)?(?:ps::)?([A-Za-z0-9_]+(?:<[^>]*>)?)", name)
    if m:
        return m.group(1).replace("ps::Fe<ps::FpParams", "Fp").replace("ps::", "")
    return name.split("(")[0].replace("void ", "")[:60]


def launches(src, dst, title):
    lines = [l for l in open(src) if not l.startswith("==")]
    agg, tot, order = collections.OrderedDict(), 0.0, []
    for r in csv.DictReader(lines):
        try:
            v = float(r["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(r["Metric Unit"], 1.0)
        k = short_name(r["Kernel Name"])
        a = agg.setdefault(k, [0, 0.0, r["Block Size"], r["Grid Size"]])
        a[0] += 1; a[1] += v; tot += v
        order.append((r["ID"], k, v))
    with open(dst, "w") as f:
        f.write("# %s\n\n" % title)
        f.write("Source: `ncu --metrics gpu__time_duration.sum --clock-control none` launch list "
                "(cold-cache, serialised: compare SHARES, not absolutes).\n\n")
        f.write("| kernel | launches | total ms | share | block | last grid |\n|---|---:|---:|---:|---|---|\n")
        for k, (c, v, blk, grd) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("| `%s` | %d | %.3f | %.1f%% | %s | %s |\n" % (k, c, v / 1e6, 100 * v / tot, blk, grd))
        # shares inside the last complete MSM step (MsmCountK ... MsmFinalK): the timed region's unit
        starts = [n for n, (_, k, _) in enumerate(order) if k.startswith("MsmCountK")]
        ends = [n for n, (_, k, _) in enumerate(order) if k.startswith("MsmFinalK")]
        if starts and ends and ends[-1] > starts[0]:
            e = ends[-1]
            b = max(x for x in starts if x < e)
            step, stot = collections.OrderedDict(), 0.0
            for _, k, v in order[b:e + 1]:
                a = step.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v; stot += v
            f.write("\n## shares inside one MSM step (launch ids %s..%s, %.3f ms in total)\n\n" % (order[b][0], order[e][0], stot / 1e6))
            f.write("| kernel | launches | total ms | share of the step |\n|---|---:|---:|---:|\n")
            for k, (c, v) in sorted(step.items(), key=lambda kv: -kv[1][1]):
                f.write("| `%s` | %d | %.3f | %.1f%% |\n" % (k, c, v / 1e6, 100 * v / stot))
        f.write("\n## launch sequence (id, kernel, us)\n\n```\n")
        for i, k, v in order:
            f.write("%s %s %.1f\n" % (i, k, v / 1e3))
        f.write("```\n")


WANT = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
    "smsp__sass_average_branch_targets_threads_uniform.pct",
]


def to_bytes(val, unit):
    v = float(val.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)


def full(rep, dst, key):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    det = subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True).stdout
    traffic = {}
    with open(dst, "w") as f:
        for vals in rows[2:]:
            name = vals[hdr.index("Kernel Name")]
            f.write("kernel: %s\n" % short_name(name))
            d = dict(zip(hdr, zip(vals, units)))
            for w in WANT:
                if w in d:
                    f.write("  %-75s %s %s\n" % (w, d[w][0], d[w][1]))
            rd, wr = to_bytes(*d["dram__bytes_read.sum"]), to_bytes(*d["dram__bytes_write.sum"])
            traffic = {"%s_dram_bytes_per_launch" % short_name(name).split("<")[0]: rd + wr, "read": rd, "write": wr}
            f.write("  dram bytes per launch (read + write): %.4g\n\n" % (rd + wr))
        f.write("---- ncu --page details ----\n")
        f.write(det)
    tj = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    cur = json.load(open(tj)) if os.path.exists(tj) else {}
    cur[key] = traffic
    json.dump(cur, open(tj, "w"), indent=1)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "launch list")
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4])
