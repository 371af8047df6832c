#!/usr/bin/env python
"""Turn ncu CSV exports into the markdown summaries kept under profiles/.

    python tools/summarize_ncu.py launches <launch-list.csv> <out.md> "<command line that was profiled>"
    python tools/summarize_ncu.py full <raw-page.csv> <out.md> "<command line>" [traffic.json kernel-regex]
"""
import collections
import csv
import json
import re
import sys

FULL_KEYS = [
    "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "sm__cycles_elapsed.max",
]


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name.replace("fpl::", "").replace("net::", "").replace("v2o::", "")


def launches(path, out, cmd):
    rows = list(csv.reader(open(path)))
    for i, r in enumerate(rows):
        if "Kernel Name" in r:
            h, start = r, i + 1
            break
    ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[start:]:
        if len(r) <= vi:
            continue
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        k = short(r[ki])
        ours = any(t in r[ki] for t in ("fpl::", "net::", "v2o::")) and not any(t in r[ki] for t in ("at::", "native::", "cudnn"))
        a = agg.setdefault(k, [0, 0.0, ours])
        a[0] += 1; a[1] += v / 1e3
    tot = sum(a[1] for a in agg.values() if a[2])
    with open(out, "w") as f:
        f.write("# launch list: `%s`\n" % cmd)
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none --csv` after the same command exited 0 "
                "without ncu.  Cold-cache, serialised times: compare SHARES (of this repo's kernels).\n"
                "torch kernels in the capture only synthesise the input volume (outside the timed region).\n\n")
        f.write("| kernel | launches | total us | share |\n|---|---|---|---|\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            if a[2]:
                f.write("| %s | %d | %.1f | %.1f %% |\n" % (k, a[0], a[1], 100 * a[1] / tot))
        other = sum(a[1] for a in agg.values() if not a[2])
        f.write("\nour kernels: %.1f us in %d launches; torch (input synthesis): %.1f us\n"
                % (tot, sum(a[0] for a in agg.values() if a[2]), other))


def full(path, out, cmd, traffic=None, kre=None):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write("# ncu --set full: `%s`\n(after the same command exited 0 without ncu)\n" % cmd)
        for vals in rows[2:]:
            d = dict(zip(hdr, vals)); u = dict(zip(hdr, units))
            f.write("\n## %s\n\n| metric | value | unit |\n|---|---|---|\n" % d.get("Kernel Name", "?"))
            for k in FULL_KEYS:
                if k in d and d[k] != "":
                    f.write("| %s | %s | %s |\n" % (k, d[k], u.get(k, "")))
            if traffic and kre and re.search(kre, d.get("Kernel Name", "")):
                def num(k):
                    v = float(d[k].replace(",", ""))
                    return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(u[k], 1.0)
                rd, wr = num("dram__bytes_read.sum"), num("dram__bytes_write.sum")
                json.dump({"dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_total": rd + wr,
                           "kernel": d["Kernel Name"], "capture": out, "note": "one launch of `%s`" % cmd},
                          open(traffic, "w"))
                traffic = None


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(*sys.argv[2:5])
    else:
        full(*sys.argv[2:7])
