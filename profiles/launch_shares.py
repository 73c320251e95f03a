"""Per-kernel shares from an `ncu --metrics gpu__time_duration.sum --csv` launch list:
    python profiles/launch_shares.py profiles/r1e_launches.csv [skip_first_n_launches]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
idx = {n: i for i, n in enumerate(rows[h])}
agg = collections.OrderedDict()
n = 0
for r in rows[h + 1:]:
    if len(r) < len(rows[h]):
        continue
    n += 1
    if n <= skip:
        continue
    val, unit = float(r[idx["Metric Value"]]), r[idx["Metric Unit"]]
    val = val / 1000.0 if unit == "ns" else (val * 1000.0 if unit == "ms" else val)
    a = agg.setdefault(r[idx["Kernel Name"]][:72], [0, 0.0])
    a[0] += 1
    a[1] += val
tot = sum(v[1] for v in agg.values())
print(f"{sys.argv[1]}: {n - skip} launches, {tot / 1000:.2f} ms (per-launch times under ncu are cold-cache and serialised: compare SHARES)")
for k, v in agg.items():
    print(f"{k:74s} n={v[0]:4d} mean={v[1] / v[0]:10.1f} us total={v[1] / 1000:8.2f} ms share={v[1] / tot:.3f}")
