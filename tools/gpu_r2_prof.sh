#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
CFG="${1:-B=49152,TAU=3}"
TAG="${2:-r2b}"
for S in 26 22; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_pr_ -s 30 -c 30 --csv \
     --log-file gpurun_out/${TAG}_launches${S}.csv python tools/pr_knobs.py --scale $S --configs "$CFG" --reps 1 > gpurun_out/${TAG}_launches${S}.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_pr_cb|k_pr_sell|k_pr_finish" -s 30 -c 3 \
     -f -o gpurun_out/${TAG}_prof${S} python tools/pr_knobs.py --scale $S --configs "$CFG" --reps 1 > gpurun_out/${TAG}_prof${S}.log 2>&1
done
grep -h "k_pr" gpurun_out/${TAG}_launches26.csv | head -12 | cut -c1-200
ls -la gpurun_out/ | tail -12
