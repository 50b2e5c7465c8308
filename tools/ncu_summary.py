#!/usr/bin/env python3
"""Summarise an `ncu --csv --metrics ...` launch list: one line per launch + totals per kernel name."""
import csv
import sys
from collections import OrderedDict, defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
idx = {h: i for i, h in enumerate(hdr)}
launches = OrderedDict()
for r in rows[1:]:
    d = launches.setdefault(r[idx['ID']], {'name': r[idx['Kernel Name']].replace('void ', '').replace('oc::', '')[:34]})
    try:
        d[r[idx['Metric Name']]] = float(r[idx['Metric Value']].replace(',', ''))
    except ValueError:
        pass
cols = [('gpu__time_duration.sum', 'us', 1e-3), ('smsp__thread_inst_executed_per_inst_executed.ratio', 'lanes', 1),
        ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue%', 1), ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps%', 1),
        ('smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio', 'noinst', 1),
        ('smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'lsb', 1),
        ('smsp__inst_executed.sum', 'Minst', 1e-6), ('dram__bytes_read.sum', 'rdMB', 1e-6), ('dram__bytes_write.sum', 'wrMB', 1e-6)]
cols = [c for c in cols if any(c[0] in d for d in launches.values())]
print(f"{'kernel':34s} " + ' '.join(f"{c[1]:>8s}" for c in cols))
tot = defaultdict(lambda: [0, 0.0])
quiet = '--totals' in sys.argv
for d in launches.values():
    if not quiet:
        print(f"{d['name']:34s} " + ' '.join(f"{d.get(c[0], 0) * c[2]:8.1f}" for c in cols))
    tot[d['name']][0] += 1
    tot[d['name']][1] += d.get('gpu__time_duration.sum', 0) * 1e-3
print('--- totals per kernel (us) ---')
for n, (c, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{n:34s} launches={c:4d} total={t:10.1f}")
print(f"all launches: {sum(t for _, t in tot.values()):.1f} us")
