#!/usr/bin/env python3
"""Per-kernel-class summary of one frame from an ncu metrics CSV:
   ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__thread_inst_executed_per_inst_executed.ratio,\\
smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum --cache-control none --clock-control none --csv \\
       --log-file X.csv python bench.py --steps 3 --warmup 2 --no-cpu-baseline
   python tools/ncu_classes.py X.csv [frame-start kernel substring, default ptGenKernel] > profiles/NAME.json
The frame taken is the one between the last two frame-start launches. Classes: GEN, TRACE, SHADE, TAIL, POST (anything else
that runs in the frame is listed under its own name). Lanes and issue utilisation are averaged with the launches' warp
instruction counts as weights. bench.py reads the JSON (measured DRAM traffic per class against its algorithmic bytes)."""
import csv
import json
import subprocess
import sys
from collections import OrderedDict

path = sys.argv[1]
start = sys.argv[2] if len(sys.argv) > 2 else "ptGenKernel"
lines = [l for l in open(path) if not l.startswith("==")]
launches = OrderedDict()
for r in csv.DictReader(lines):
    k = int(r["ID"])
    d = launches.setdefault(k, {"name": r["Kernel Name"]})
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"]
    m = r["Metric Name"]
    if m == "gpu__time_duration.sum":
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1e-6)
    if m.startswith("dram__bytes"):
        v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
    d[m] = v
ids = list(launches)
starts = [i for i in ids if start in launches[i]["name"]]
if len(starts) < 2:
    raise SystemExit(f"need two '{start}' launches in the capture, found {len(starts)}")
frame = [launches[i] for i in ids if starts[-2] <= i < starts[-1]]


def cls(name):
    for key, c in (("ptGenKernel", "GEN"), ("aoGenKernel", "GEN"), ("ptTraceKernel", "TRACE"), ("ptShadeKernel", "SHADE"),
                   ("aoShadeKernel", "SHADE"), ("ptTailKernel", "TAIL"), ("postprocessKernel", "POST"), ("renderFrameKernel", "RENDER")):
        if key in name:
            return c
    return None


out = OrderedDict()
for l in frame:
    c = cls(l["name"])
    if c is None:
        continue
    o = out.setdefault(c, {"launches": 0, "ms": 0.0, "dram_read_bytes": 0.0, "dram_write_bytes": 0.0, "warp_inst": 0.0, "_lanes": 0.0, "_issue": 0.0})
    wi = l.get("smsp__inst_executed.sum", 0.0)
    o["launches"] += 1
    o["ms"] += l.get("gpu__time_duration.sum", 0.0)
    o["dram_read_bytes"] += l.get("dram__bytes_read.sum", 0.0)
    o["dram_write_bytes"] += l.get("dram__bytes_write.sum", 0.0)
    o["warp_inst"] += wi
    o["_lanes"] += wi * l.get("smsp__thread_inst_executed_per_inst_executed.ratio", 0.0)
    o["_issue"] += l.get("gpu__time_duration.sum", 0.0) * l.get("smsp__issue_active.avg.pct_of_peak_sustained_active", 0.0)
for c, o in out.items():
    o["lanes_per_instruction"] = round(o.pop("_lanes") / o["warp_inst"], 2) if o["warp_inst"] else None
    o["issue_slot_utilisation_pct"] = round(o.pop("_issue") / o["ms"], 1) if o["ms"] else None
    o["ms"] = round(o["ms"], 4)
    o["dram_bytes"] = o["dram_read_bytes"] + o["dram_write_bytes"]
try:
    commit = subprocess.check_output(["git", "rev-parse", "--short", "HEAD"], text=True).strip()
except Exception:
    commit = None
json.dump({"source": path, "commit": commit, "how": "ncu per-launch metrics, warm caches (--cache-control none), one frame; times are serialised "
           "under the profiler: use the SHARE per class, not the absolute", "classes": out}, sys.stdout, indent=1)
print()
