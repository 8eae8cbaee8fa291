#!/bin/bash
# per-kernel device times of a few sweeps (ncu serialises the kernels: compare shares)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TAG="${1:-r2f}"; CFG="${2:-B=32768,TAU=2,DUAL=0}"; SCALES="${3:-26 22}"
for S in $SCALES; do
  N=$(echo $CFG | tr -c 'A-Za-z0-9' '_')
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_pr_ -s 40 -c 12 --csv \
     --log-file gpurun_out/${TAG}_launches${S}_${N}.csv python tools/pr_knobs.py --scale $S --configs "$CFG" --reps 1 > gpurun_out/${TAG}_launches${S}_${N}.log 2>&1
  echo "scale $S $CFG"; grep -h "k_pr" gpurun_out/${TAG}_launches${S}_${N}.csv | head -8 | awk -F'","' '{print $5, $(NF-6), $(NF-5), $NF}'
  tail -2 gpurun_out/${TAG}_launches${S}_${N}.log | cut -c1-900
done
