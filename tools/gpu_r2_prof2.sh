#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TAG="${1:-r2d}"
for CFG in "B=32768,TAU=2,DUAL=0" "B=32768,TAU=2,DUAL=1"; do
  N=$(echo $CFG | tr -c 'A-Za-z0-9' '_')
  for S in 26 22; do
    timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_pr_ -s 40 -c 16 --csv \
       --log-file gpurun_out/${TAG}_launches${S}_${N}.csv python tools/pr_knobs.py --scale $S --configs "$CFG" --reps 1 > gpurun_out/${TAG}_launches${S}_${N}.log 2>&1
    grep -h "k_pr" gpurun_out/${TAG}_launches${S}_${N}.csv | head -8 | awk -F'","' '{print $5, $NF}'
  done
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_pr_cb|k_pr_sell|k_pr_finish" -s 40 -c 4 \
   -f -o gpurun_out/${TAG}_prof26_dual python tools/pr_knobs.py --scale 26 --configs "B=32768,TAU=2,DUAL=1" --reps 1 > gpurun_out/${TAG}_prof26.log 2>&1
timeout 300 python tools/pr_knobs.py --scale 26 --configs "B=32768,TAU=2,DUAL=0;B=24576,TAU=2;B=16384,TAU=2;B=24576,TAU=3;B=32768,TAU=2,HOT=0" > gpurun_out/${TAG}_knobs26.jsonl 2> gpurun_out/${TAG}_knobs26.err
cut -c1-120 gpurun_out/${TAG}_knobs26.jsonl
timeout 300 python tools/pr_knobs.py --scale 22 --configs "B=32768,TAU=2,DUAL=0;B=24576,TAU=2;B=16384,TAU=2;B=16384,TAU=2,DUAL=0" > gpurun_out/${TAG}_knobs22.jsonl 2> gpurun_out/${TAG}_knobs22.err
cut -c1-120 gpurun_out/${TAG}_knobs22.jsonl
