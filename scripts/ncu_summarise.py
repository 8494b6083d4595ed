"""Summarise a `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, total device
time and share of the step.  Usage: python scripts/ncu_summarise.py gpurun_out/ncu_<tag>_launches.csv"""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
hdr = rows[hi]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value'); ui = hdr.index('Metric Unit')
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hi + 1:]:
    if len(r) <= vi:
        continue
    name = re.sub(r'\(.*', '', r[ki]); name = re.sub(r'^void ', '', name)
    v = float(r[vi].replace(',', '')); u = r[ui]
    v *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6}.get(u, 1.0)
    agg[name][0] += 1; agg[name][1] += v
tot = sum(v[1] for v in agg.values())
print(f"# {sys.argv[1]}: {sum(v[0] for v in agg.values())} launches, {tot / 1e3:.1f} ms of kernel time (cold-cache, serialised)")
print(f"{'kernel':70s} {'launches':>8s} {'total ms':>10s} {'share':>7s} {'avg us':>10s}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:70]:70s} {v[0]:8d} {v[1] / 1e3:10.2f} {v[1] / tot * 100:6.2f}% {v[1] / v[0]:10.1f}")
