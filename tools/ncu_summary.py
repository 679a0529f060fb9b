"""Developer tool: one line per kernel from `ncu -i X.ncu-rep --page raw --csv` (stdin or a file), plus an optional traffic JSON
(dram bytes per launch / per read) that bench.py's roofline.traffic reads.
usage: ncu -i rep.ncu-rep --page raw --csv > raw.csv; python tools/ncu_summary.py raw.csv [n_reads traffic.json tag]"""
import csv
import json
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
head, units, data = rows[0], rows[1], rows[2:]
col = {n: i for i, n in enumerate(head)}


def val(r, name, want=None):
    i = col.get(name)
    if i is None or r[i] in ("", "no data"):
        return None
    v = float(r[i].replace(",", ""))
    u = units[i]
    scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "byte": 1.0, "us": 1e-3, "ms": 1.0, "s": 1e3, "ns": 1e-6}
    return v * scale.get(u, 1.0)


out = {}
for r in data:
    name = re.sub(r"\(.*", "", r[col["Kernel Name"]]).split("<")[0]
    ms = val(r, "gpu__time_duration.sum")
    rd, wr = val(r, "dram__bytes_read.sum") or 0.0, val(r, "dram__bytes_write.sum") or 0.0
    req, sec = val(r, "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum"), val(r, "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum")
    line = (f"{name:<18} {ms:7.3f} ms  dram rd {rd / 1e9:6.2f} GB wr {wr / 1e9:6.2f} GB -> {(rd + wr) / 1e9 / (ms * 1e-3):7.1f} GB/s"
            f"  L2 hit {val(r, 'lts__t_sector_hit_rate.pct') or 0:5.1f}%  L1 hit {val(r, 'l1tex__t_sector_hit_rate.pct') or 0:5.1f}%"
            f"  sectors/request(ld) {(sec / req) if req else 0:5.2f}"
            f"  issue {val(r, 'sm__issue_active.avg.pct_of_peak_sustained_elapsed') or 0:5.1f}%"
            f"  warps {val(r, 'sm__warps_active.avg.pct_of_peak_sustained_active') or 0:5.1f}%"
            f"  regs {int(val(r, 'launch__registers_per_thread') or 0)}"
            f"  stall/issue: long_sb {val(r, 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio') or 0:.2f}"
            f" no_inst {val(r, 'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio') or 0:.2f}"
            f" math {val(r, 'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio') or 0:.2f}"
            f" wait {val(r, 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio') or 0:.2f}")
    print(line)
    if name not in out:      # the first launch of a kernel is the primary pass; later ones (re-map, big arena) are small
        out[name] = {"dram_bytes_per_launch": rd + wr, "ms_under_ncu": ms}
if len(sys.argv) > 4:
    n_reads, path, tag = int(sys.argv[2]), sys.argv[3], sys.argv[4]
    try:
        tj = json.load(open(path))
    except Exception:
        tj = {}
    for k, v in out.items():
        tj[k] = {"dram_bytes_per_launch": v["dram_bytes_per_launch"], "dram_bytes_per_read": v["dram_bytes_per_launch"] / n_reads,
                 "ms_under_ncu": v["ms_under_ncu"], "from": tag}
    json.dump(tj, open(path, "w"), indent=1)
