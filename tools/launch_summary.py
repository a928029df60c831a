"""Per-kernel time shares of the assembly steps from an `ncu --metrics gpu__time_duration.sum --csv` launch list.
usage: python tools/launch_summary.py launches.csv [n_steps]   (the last n_steps occurrences of the fused kernel
delimit the steps; everything before is plan build / warm-up)"""
import csv
import sys
from collections import OrderedDict

path = sys.argv[1]
rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
names = [r[4] for r in rows]
times = [float(r[-1].replace(",", "")) for r in rows]
units = rows[0][-2] if rows else "?"


def short(n):
    n = n.replace("<unnamed>::", "")
    cut = n.find("(")
    n = n if cut < 0 else n[:cut]
    return n[:90]


fused = [i for i, n in enumerate(names) if "fused_h1_kernel" in n or "p2tet_elasticity_kernel" in n]
print("launches:", len(rows), "unit:", units, "fused launches:", len(fused))
if len(fused) >= 2:
    # one step = from the geo_records launch preceding a fused launch up to the next one
    starts = [i - 1 for i in fused]
    last = starts[-2], starts[-1]
    step = OrderedDict()
    for i in range(last[0], last[1]):
        step[short(names[i])] = step.get(short(names[i]), 0.0) + times[i]
    total = sum(step.values())
    print("one steady-state step (device time, serialised, cold caches): %.1f %s" % (total, units))
    for k, v in step.items():
        print("  %6.1f%%  %10.1f  %s" % (100 * v / total, v, k))
build = OrderedDict()
for i in range(0, fused[0] - 1 if fused else len(rows)):
    build[short(names[i])] = build.get(short(names[i]), 0.0) + times[i]
tot = sum(build.values())
print("before the first step (coefficient mapping, plan build): %.1f %s" % (tot, units))
for k, v in sorted(build.items(), key=lambda kv: -kv[1])[:12]:
    print("  %6.1f%%  %10.1f  %s" % (100 * v / tot, v, k))
