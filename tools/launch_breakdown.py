"""Development aid: per-kernel time of ONE build taken from an ncu launch list (gpu__time_duration.sum, --csv)."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
which = int(sys.argv[2]) if len(sys.argv) > 2 else 1
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H = rows[hdr]
ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
names = []
for r in rows[hdr + 1:]:
    if len(r) <= vi:
        continue
    n = r[ki].split("(")[0].replace("atlas::<unnamed>::", "").replace("void ", "")
    names.append((n, float(r[vi].replace(",", "")) * (1e-3 if r[ui] == "ns" else 1.0)))
starts = [i for i, (n, v) in enumerate(names) if n == "init_refs"] + [len(names)]
seg = names[starts[which]:starts[which + 1]]
agg = collections.OrderedDict()
for n, v in seg:
    a = agg.setdefault(n, [0, 0.0])
    a[0] += 1
    a[1] += v
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{c:4d} {t:9.1f} us {t / c:8.1f} avg  {k}")
print(f"{sum(t for _, t in agg.values()):.1f} us in {len(seg)} launches")
for k in ("bin_big", "partition_scatter", "select_big"):
    print(k, [round(v, 1) for n, v in seg if n == k])
