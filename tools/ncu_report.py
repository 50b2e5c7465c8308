#!/usr/bin/env python3
"""Text summary of an `ncu --set full` capture for profiles/:
   python tools/ncu_report.py X.ncu-rep [hot-line count] > profiles/NAME.txt
Per captured launch: duration, launch shape, occupancy, lanes per instruction, issue utilisation, stall reasons,
cache hit rates, DRAM bytes, pipe utilisation; then the hottest source lines (needs -lineinfo + --import-source on)."""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
nhot = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
want = [
    ("gpu__time_duration.sum", "duration"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
    ("launch__registers_per_thread", "registers/thread"), ("launch__occupancy_limit_registers", "occupancy limit (regs), CTAs/SM"),
    ("launch__occupancy_limit_shared_mem", "occupancy limit (smem), CTAs/SM"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy (warps active %)"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active lanes per instruction (of 32)"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slot utilisation %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed", "memory throughput %"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"), ("lts__t_bytes.sum", "L2 bytes"),
    ("smsp__inst_executed_op_local_ld.sum", "local loads"), ("smsp__inst_executed_op_local_st.sum", "local stores"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared bank conflicts"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "pipe alu %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "pipe fma %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "pipe lsu %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "pipe xu %"),
]
print(f"source: {rep.split('/')[-1]} (ncu --set full --clock-control none --import-source on)")
for r in rows[2:]:
    print(f"\n== {r[col['Kernel Name']]}  (launch id {r[col['ID']]})")
    for m, label in want:
        if m in col:
            print(f"  {label:42s} {r[col[m]]:>16s} {units[col[m]]}")
    st = []
    for h in hdr:
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and "not_issued" not in h:
            try:
                st.append((float(r[col[h]].replace(",", "")), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
            except ValueError:
                pass
    print("  stalled warps per issued instruction: " + ", ".join(f"{n} {v:.2f}" for v, n in sorted(st, reverse=True)[:7]))

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur_file, h2, agg, cur_line = None, None, {}, None
for r in csv.reader(io.StringIO(src)):
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        h2 = {h: i for i, h in enumerate(r)}
        continue
    if h2 is None or len(r) < 10:
        continue
    if r[0].strip().isdigit():
        cur_line = (int(r[0]), r[1].strip()[:100])
        continue
    if r[0].strip() != "" or cur_line is None:
        continue
    try:
        inst, tinst, samp = float(r[h2["Instructions Executed"]]), float(r[h2["Thread Instructions Executed"]]), float(r[h2["# Samples"]])
    except (ValueError, KeyError):
        continue
    if inst == 0 and samp == 0:
        continue
    a = agg.setdefault((cur_file, cur_line[0]), [0.0, 0.0, 0.0, cur_line[1]])
    a[0] += inst
    a[1] += tinst
    a[2] += samp
if agg:
    tot = sum(a[0] for a in agg.values()) or 1
    tots = sum(a[2] for a in agg.values()) or 1
    print(f"\nhottest source lines (all captured launches; share of warp instructions, of stall samples, active lanes):")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][2])[:nhot]:
        lanes = a[1] / a[0] if a[0] else 0
        print(f"  {k[0]:20s}:{k[1]:4d} inst {100*a[0]/tot:5.1f}% samples {100*a[2]/tots:5.1f}% lanes {lanes:4.1f} | {a[3]}")
