#!/usr/bin/env python
"""Turn an .ncu-rep (one kernel, --set full) into the short markdown summary kept under profiles/."""
import csv
import subprocess
import sys

rep, title = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
m = {h: (u, v) for h, u, v in zip(hdr, units, vals)}


def g(k):
    return m.get(k, ("", "n/a"))


want = [
    ("Kernel Name", "kernel"), ("launch__grid_size", "grid (CTAs)"), ("launch__block_size", "block (threads)"),
    ("launch__registers_per_thread", "registers / thread"), ("launch__waves_per_multiprocessor", "waves / SM"),
    ("gpu__time_duration.sum", "duration (cold, under ncu)"),
    ("dram__bytes_read.sum", "DRAM bytes read"), ("dram__bytes_write.sum", "DRAM bytes written"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"), ("l1tex__t_sector_hit_rate.pct", "L1 hit rate %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe busy %"),
    ("sm__inst_executed_pipe_fp64.sum", "FP64 warp instructions"),
    ("smsp__inst_executed.sum", "warp instructions executed"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared-memory / shuffle wavefronts"),
    ("sm__cycles_elapsed.max", "SM cycles elapsed"),
]
print(f"# {title}\n")
print(f"source: `{rep.split('/')[-1]}` (ncu --set full --clock-control none --import-source on; one launch)\n")
print("| metric | value |\n|---|---|")
for k, label in want:
    u, v = g(k)
    print(f"| {label} | {v} {u} |")
stalls = sorted(((float(v), h) for h, (u, v) in m.items()
                 if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")
                 and "not_issued" not in h and v not in ("", "n/a")), reverse=True)
print("\nWarp stall cycles per issued instruction (top reasons):\n")
print("| reason | cycles / issue |\n|---|---|")
for v, h in stalls[:8]:
    print(f"| {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')} | {v:.2f} |")
