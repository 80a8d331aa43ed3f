"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: launches, total and mean device time per kernel of this library."""
import csv
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
agg = OrderedDict()
for r in rows:
    name = r[4]
    short = name.split("(")[0].replace("void ", "").replace("pstb::", "").replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
    if "at::" in name or "cub::" in name or "distribution" in name or "elementwise" in name:
        short = "[torch] " + short[:60]
    a = agg.setdefault(short, [0, 0.0, r[7], r[8]])
    a[0] += 1
    a[1] += float(r[14].replace(",", ""))
ours = {k: v for k, v in agg.items() if not k.startswith("[torch]")}
tot = sum(v[1] for v in ours.values()) or 1.0
print("# %d launches captured, %d of this library's kernels" % (len(rows), sum(v[0] for v in ours.values())))
for k, v in sorted(ours.items(), key=lambda kv: -kv[1][1]):
    print("%6d launches %12.3f ms total %10.3f ms/launch %5.1f %%  %s  grid %s block %s" % (v[0], v[1] / 1e6, v[1] / 1e6 / v[0], 100 * v[1] / tot, k[:90], v[3], v[2]))
t = [v for k, v in agg.items() if k.startswith("[torch]")]
print("# torch (synthetic store generation, copies): %d launches, %.3f ms" % (sum(v[0] for v in t), sum(v[1] for v in t) / 1e6))
