"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): total ms and launch count per kernel."""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
h = rows[hi]; idx = {n: i for i, n in enumerate(h)}
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hi + 1:]:
    if len(r) < len(h) or r[idx['Metric Name']] != 'gpu__time_duration.sum': continue
    v = float(r[idx['Metric Value']].replace(',', '')); u = r[idx['Metric Unit']]
    v *= {'ns': 1e-6, 'nsecond': 1e-6, 'us': 1e-3, 'usecond': 1e-3, 'ms': 1.0, 'msecond': 1.0}[u]
    k = r[idx['Kernel Name']][:int(sys.argv[2]) if len(sys.argv) > 2 else 80]
    agg[k][0] += 1; agg[k][1] += v
tot = sum(t for _, t in agg.values())
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]): print(f"{t:10.3f} ms {100*t/tot:5.1f}% {n:6d}  {k}")
print(f"{tot:10.3f} ms total")
