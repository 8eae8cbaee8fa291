#!/usr/bin/env python3
"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections, csv, sys
agg = collections.defaultdict(lambda: [0, 0.0])
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
for row in csv.DictReader(lines):
    v = float(row['Metric Value'].replace(',', ''))
    unit = row['Metric Unit']
    v = v / 1e3 if unit == 'ns' else (v * 1e3 if unit == 'ms' else v)
    k = row['Kernel Name'][:70]
    agg[k][0] += 1
    agg[k][1] += v
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 20]:
    print(f"{k:72s} n={v[0]:5d} total={v[1]:11.1f}us mean={v[1]/v[0]:10.1f}us share={v[1]/tot:.3f}")
