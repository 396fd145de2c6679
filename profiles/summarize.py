#!/usr/bin/env python
"""Summarise ncu captures (gpurun_out/*.ncu-rep + launch list) into the tracked profiles/ files.

    python profiles/summarize.py r2 C2x300 gpurun_out/r2_launches.csv gpurun_out/r2_prof_*.ncu-rep

`C2x300` names the workload the captures ran (bench.py copies `traffic` only when it runs the same one).

Writes profiles/<round>_launches_summary.csv, profiles/<round>_kernels.csv, profiles/<round>_summary.md and
profiles/latest.json (dram bytes per launch of the dominant kernel: bench.py's roofline.traffic)."""
import collections
import csv
import io
import json
import os
import statistics
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall branch"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg_throttle"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall no_instruction"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not_selected"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe"),
]


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def raw_rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    res = []
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("(")[0]
        for junk in ("void ", "b2v::"):
            name = name.replace(junk, "")
        d = {"kernel": name.split("<")[0]}
        for m, name in METRICS:
            if m in idx:
                d[name] = (r[idx[m]], units[idx[m]])
        res.append(d)
    return res


def main():
    rnd, workload, launches, reps = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4:]
    # ---- launch list: per-kernel totals and shares (cold-cache, serialised: compare SHARES) ----
    rows = [r for r in csv.reader(open(launches)) if len(r) > 10]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    per = collections.defaultdict(list)
    for r in rows[1:]:
        try:
            per[r[ki].split("(")[0].replace("void ", "").replace("b2v::", "")].append(float(r[vi].replace(",", "")))
        except ValueError:
            pass
    total = sum(sum(v) for v in per.values())
    with open(os.path.join(HERE, f"{rnd}_launches_summary.csv"), "w") as f:
        f.write("kernel,launches,mean_us,median_us,total_ms,share\n")
        for k, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"{k},{len(v)},{statistics.mean(v) / 1e3:.2f},{statistics.median(v) / 1e3:.2f},"
                    f"{sum(v) / 1e6:.3f},{sum(v) / total:.4f}\n")
    # ---- full captures ----
    kernels = []
    for rep in reps:
        kernels += raw_rows(rep)
    names = [n for _, n in METRICS]
    with open(os.path.join(HERE, f"{rnd}_kernels.csv"), "w") as f:
        f.write("kernel," + ",".join(names) + "\n")
        for d in kernels:
            f.write(d["kernel"] + "," + ",".join(f"{d.get(n, ('', ''))[0]} {d.get(n, ('', ''))[1]}".strip() for n in names) + "\n")
    latest = {}
    by = collections.defaultdict(list)
    for d in kernels:
        by[d["kernel"]].append(d)
    for k, ds in by.items():
        rd = statistics.mean(to_bytes(*d["dram read"]) for d in ds)
        wr = statistics.mean(to_bytes(*d["dram write"]) for d in ds)
        latest[k] = {"dram_bytes_per_launch": rd + wr, "dram_read_bytes": rd, "dram_write_bytes": wr,
                     "duration_us": statistics.mean(float(d["duration"][0]) for d in ds), "captures": len(ds),
                     "workload": workload, "source": "ncu --set full, profiles/" + rnd + "_kernels.csv"}
        if "issue active %" in ds[0]:
            latest[k]["issue_active_pct"] = statistics.mean(float(d["issue active %"][0]) for d in ds)
    latest["round"] = rnd
    latest["source"] = [os.path.basename(r) for r in reps]
    json.dump(latest, open(os.path.join(HERE, "latest.json"), "w"), indent=1)
    print(json.dumps(latest, indent=1))


if __name__ == "__main__":
    main()
