"""Per-kernel launch counts and mean / min duration from an `ncu --metrics gpu__time_duration.sum --csv` launch list.
usage: python tools/ncu_launch_summary.py <launches.csv>"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr = rows[hi]
kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= mv:
        continue
    v = float(r[mv].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[mu], 1.0)
    agg.setdefault(r[kn][:100], []).append(v)
tot = sum(sum(v) for v in agg.values())
print(f"{'kernel':100s} {'n':>4s} {'mean ms':>9s} {'min ms':>9s} {'share':>6s}")
for k, v in agg.items():
    print(f"{k:100s} {len(v):4d} {sum(v) / len(v):9.4f} {min(v):9.4f} {sum(v) / tot:6.1%}")
