#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into markdown: duration, DRAM/L2 traffic, throughput %, occupancy,
stall mix, tensor-pipe activity. Usage: ncu_summarize.py report.ncu-rep [title] > profiles/x.md"""
import csv
import subprocess
import sys

rep = sys.argv[1]
title = sys.argv[2] if len(sys.argv) > 2 else rep
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"), ("launch__shared_mem_per_block_static", "static smem/block"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__t_bytes.sum", "L2 bytes"), ("lts__t_sector_hit_rate.pct", "L2 hit rate"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit rate"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput % of peak"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("sm__pipe_tensor_subpipe_ummа_cycles_active.avg.pct_of_peak_sustained_active", "tensor umma active %"),
    ("sm__inst_executed_pipe_uniform.sum", "uniform-pipe insts"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__cycles_elapsed.max", "SM cycles elapsed"),
    ("lts__t_sectors_srcunit_tex_aperture_peer.sum", "L2 sectors to/from peer aperture"),
]
STALLS = [k for k in hdr if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")]
print(f"# {title}\n")
print(f"source: `{rep}` (ncu --set full --clock-control none --import-source on)\n")
for n, r in enumerate(data):
    print(f"## launch {n}: `{r[idx['Kernel Name']][:110]}`\n")
    print("| metric | value |\n|---|---|")
    for k, label in KEYS:
        if k in idx and r[idx[k]] != "":
            print(f"| {label} | {r[idx[k]]} {units[idx[k]]} |")
    tens = [k for k in hdr if "tensor" in k and "pct" in k]
    for k in tens[:6]:
        if r[idx[k]] not in ("", "0"):
            print(f"| {k} | {r[idx[k]]} {units[idx[k]]} |")
    st = sorted(((float(r[idx[k]] or 0), k) for k in STALLS), reverse=True)[:6]
    print("\nTop stall reasons (warps stalled per issue-active cycle):\n")
    for v, k in st:
        print(f"* {k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')}: {v:.2f}")
    print()
