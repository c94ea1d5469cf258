"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total, mean, share."""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("=="))]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    if len(r) <= vi:
        continue
    name = re.sub(r"\(.*", "", r[ki])
    v = float(r[vi].replace(",", ""))
    u = r[ui]
    ms = v / 1e6 if u in ("ns", "nsecond") else v / 1e3 if u in ("us", "usecond") else v if u in ("ms", "msecond") else v * 1e3
    tot[name][0] += 1
    tot[name][1] += ms
ours = {k: v for k, v in tot.items() if "at::" not in k and "cutlass" not in k and "gemm" not in k.lower() and "nccl" not in k.lower() and "elementwise" not in k}
total = sum(v[1] for v in ours.values())
print("# shares of device time per kernel of this library (torch fills / the int8 GEMM peak measurement excluded); %s" % path)
for k, (n, ms) in sorted(ours.items(), key=lambda kv: -kv[1][1]):
    print("%-70s n=%3d total_ms=%9.3f avg_ms=%8.4f share=%6.2f%%" % (k[:70], n, ms, ms / n, 100 * ms / total))
