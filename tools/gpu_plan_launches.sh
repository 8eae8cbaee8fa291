#!/bin/bash
# per-kernel times of the layout build (ncu launch list of tools/plan_time.py, one rep)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TAG="${1:-plan}"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_cb|k_mega|k_sell|k_perm|k_lens|k_fill|k_blk|k_rows|k_count|k_loc|k_gather|Scan|Radix|k_feed" \
  --csv --log-file gpurun_out/${TAG}_launches.csv python tools/plan_time.py --scale 26 --reps 1 > gpurun_out/${TAG}_ncu.log 2>&1
python - <<PY
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/${TAG}_launches.csv")) if len(r) > 5]
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
H = rows[hdr]; ki, vi = H.index("Kernel Name"), H.index("Metric Value")
ui = H.index("Metric Unit")
acc = collections.OrderedDict()
for r in rows[hdr + 1:]:
    v = float(r[vi].replace(",", ""))
    if r[ui] in ("ns", "nsecond"): v /= 1e6
    elif r[ui] in ("us", "usecond"): v /= 1e3
    k = r[ki][:60]
    acc.setdefault(k, [0, 0.0]); acc[k][0] += 1; acc[k][1] += v
for k, (c, v) in sorted(acc.items(), key=lambda kv: -kv[1][1])[:24]:
    print(f"{v:9.3f} ms  x{c:<4d} {k}")
PY
