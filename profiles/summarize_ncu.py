#!/usr/bin/env python
"""Summarise an `ncu --set full` report of one bench step into the two artefacts bench.py and DESIGN.md cite:

    python profiles/summarize_ncu.py gpurun_out/full_v4.ncu-rep 32768 profiles/r1_ncu_full_v4

writes <prefix>_raw.csv (ncu --page raw --csv), <prefix>_summary.txt (one line per launch) and
profiles/r1_ncu_traffic.json (dram__bytes_read.sum + dram__bytes_write.sum per launch and per read of the LARGEST launch
of each kernel, i.e. the primary pass)."""
import csv
import json
import os
import subprocess
import sys

rep, n_reads, prefix = sys.argv[1], int(sys.argv[2]), sys.argv[3]
traffic_name = sys.argv[4] if len(sys.argv) > 4 else "r1_ncu_traffic.json"
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
open(prefix + "_raw.csv", "w").write(raw)
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}


def val(r, name, unit_scale=None):
    v = float(r[ix[name]] or 0)
    u = units[ix[name]]
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(u, 1.0)
    return v * scale


lines, traffic = [], {}
for r in rows[2:]:
    name = r[ix["Kernel Name"]].split("(")[0]
    ms = val(r, "gpu__time_duration.sum")
    rd, wr = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum")
    line = (f"{name:16s} {ms:8.3f} ms  dram rd {rd / 1e9:6.2f} GB wr {wr / 1e9:6.2f} GB -> {(rd + wr) / ms / 1e6:7.1f} GB/s"
            f"  L2 hit {float(r[ix['lts__t_sector_hit_rate.pct']] or 0):5.1f}%  L1 hit {float(r[ix['l1tex__t_sector_hit_rate.pct']] or 0):5.1f}%"
            f"  issue {float(r[ix['smsp__issue_active.avg.pct_of_peak_sustained_active']] or 0):5.1f}%"
            f"  warps {float(r[ix['sm__warps_active.avg.pct_of_peak_sustained_active']] or 0):5.1f}%"
            f"  regs {r[ix['launch__registers_per_thread']]}"
            f"  thr/inst {float(r[ix['smsp__thread_inst_executed_per_inst_executed.ratio']] or 0):5.2f}"
            f"  stalls: no_inst {r[ix['smsp__pcsamp_warps_issue_stalled_no_instructions']]} long_sb {r[ix['smsp__pcsamp_warps_issue_stalled_long_scoreboard']]}"
            f" wait {r[ix['smsp__pcsamp_warps_issue_stalled_wait']]} selected {r[ix['smsp__pcsamp_warps_issue_stalled_selected']]}")
    lines.append(line)
    if name not in traffic or ms > traffic[name]["ms_under_ncu"]:
        traffic[name] = {"dram_bytes_per_launch": rd + wr, "dram_bytes_per_read": (rd + wr) / n_reads, "ms_under_ncu": ms}
open(prefix + "_summary.txt", "w").write("\n".join(lines) + "\n")
json.dump(traffic, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), traffic_name), "w"), indent=1)
print("\n".join(lines))
