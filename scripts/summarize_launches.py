"""ncu launch-list CSV (gpu__time_duration.sum per launch) -> per-kernel table: launches, total us, share of the step."""
import csv, sys
from collections import defaultdict
rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
kn, mv = rows[h].index("Kernel Name"), rows[h].index("Metric Value")
d = defaultdict(lambda: [0, 0.0])
for r in rows[h + 1:]:
    if len(r) > mv:
        d[r[kn][:90]][0] += 1
        d[r[kn][:90]][1] += float(r[mv].replace(",", ""))
tot = sum(v[1] for v in d.values())
print(f"total GPU time of the profiled step (cold-cache, serialised by ncu): {tot / 1e6:.3f} ms over {sum(v[0] for v in d.values())} launches")
for k, v in sorted(d.items(), key=lambda kv: -kv[1][1]):
    print(f"{v[0]:4d} x  {v[1] / 1e3:10.1f} us  {100 * v[1] / tot:5.1f} %  {k}")
