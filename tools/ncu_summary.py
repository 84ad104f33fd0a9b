#!/usr/bin/env python
"""Writes profiles/<tag>_ncu_summary.md from gpurun_out/prof_<tag>.ncu-rep and gpurun_out/launches_<tag>.csv."""
import csv
import subprocess
import sys
from collections import defaultdict

tag = sys.argv[1]
rep = f"gpurun_out/prof_{tag}.ncu-rep"
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(raw))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]
body = rows[2:]
names = [r[hdr.index("Kernel Name")].split("(")[0] for r in body]
out = [f"# {tag}: `ncu --set full --clock-control none` of the two heaviest kernels\n\n",
       f"Command: `bash tools/gpu_profile.sh {tag}` (bench.py default workload: 38760 frames of C2; one launch of each "
       "kernel, cold cache, serialised -- compare shares, not absolutes).  The .ncu-rep stays in gpurun_out/ (scratch).\n\n",
       "| metric | " + " | ".join(names) + " |\n", "|---|" + "---|" * len(names) + "\n"]
for k in want:
    if k in hdr:
        i = hdr.index(k)
        out.append(f"| {k} [{units[i]}] | " + " | ".join(r[i] for r in body) + " |\n")
# launch list
try:
    lr = list(csv.reader(open(f"gpurun_out/launches_{tag}.csv")))
    hi = next(i for i, r in enumerate(lr) if r and r[0] == "ID")
    h = lr[hi]
    kn, mv = h.index("Kernel Name"), h.index("Metric Value")
    agg = defaultdict(lambda: [0, 0.0])
    for r in lr[hi + 1:]:
        if len(r) > mv:
            name = r[kn].split("(")[0]
            agg[name][0] += 1
            agg[name][1] += float(r[mv].replace(",", ""))
    tot = sum(v[1] for v in agg.values())
    unit = lr[hi + 1][h.index("Metric Unit")]
    out.append(f"\n## Launch list (`--metrics gpu__time_duration.sum`, first 80 launches of the same command)\n\n"
               f"| kernel | launches | total [{unit}] | share |\n|---|---|---|---|\n")
    for name, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| {name} | {v[0]} | {v[1]:.1f} | {100 * v[1] / tot:.1f} % |\n")
except Exception as e:  # noqa: BLE001
    out.append(f"\n(launch list unavailable: {e})\n")
open(f"profiles/{tag}_ncu_summary.md", "w").write("".join(out))

# machine-readable numbers for bench.py (roofline.traffic, issue_roofline): first launch of each heavy kernel
import json
roles = {"fb_k0_ingest": "ingest", "fb_k1_analyze": "analyze", "fb_ka_plan": "plan", "fb_kp_pack": "pack"}
col = {k: hdr.index(k) for k in ("launch__grid_size", "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
                                  "gpu__time_duration.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active") if k in hdr}


def num(r, k):
    v = float(r[col[k]].replace(",", ""))
    u = units[col[k]]
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1.0, "us": 1e-3, "ns": 1e-6}.get(u, 1.0)
    return v * scale


frames = None
for r, name in zip(body, names):
    if name.startswith("fb_ka_plan") or name.startswith("fb_kp_pack"):
        frames = int(num(r, "launch__grid_size"))
        break
kern = {}
for r, name in zip(body, names):
    role = next((v for k, v in roles.items() if name.startswith(k)), None)
    if role is None or role in kern:
        continue
    kern[role] = {"kernel": name, "frames": frames, "warp_inst": num(r, "smsp__inst_executed.sum"),
                  "dram_bytes": num(r, "dram__bytes_read.sum") + num(r, "dram__bytes_write.sum"),
                  "time_ms": num(r, "gpu__time_duration.sum"),
                  "issue_active_pct": num(r, "smsp__issue_active.avg.pct_of_peak_sustained_active")}
json.dump({"tag": tag, "source": f"profiles/{tag}_ncu_summary.md (ncu --set full --clock-control none, first launch of each kernel "
                                 "of `python bench.py --steps 1 --warmup 1`)", "kernels": kern},
          open("profiles/ncu_metrics.json", "w"), indent=1)
print("".join(out))
