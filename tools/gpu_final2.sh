#!/bin/bash
# final single-GPU pass without the test suite (run separately): smoke, bench lines, ncu launch list + traffic
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TAG="${1:-r2z}"
timeout 300 python -c "import __graft_entry__ as e; e.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -1 gpurun_out/${TAG}_smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench26.json 2> gpurun_out/${TAG}_bench26.err
python -c "
import json; d=json.load(open('gpurun_out/${TAG}_bench26.json')); print('s26', d['value'], d.get('verified'), d['roofline']['frac'], d['e2e']['value'], d['e2e']['ms_per_step'], d['cpu_baseline'], d['gpu_launches'])"
timeout 600 python bench.py --steps 5 --warmup 3 --scale 22 --no-cpu > gpurun_out/${TAG}_bench22.json 2> gpurun_out/${TAG}_bench22.err
python -c "
import json; d=json.load(open('gpurun_out/${TAG}_bench22.json')); print('s22', d['value'], d.get('verified'), d['roofline']['frac'], d['e2e']['value'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_pr_ -s 60 -c 60 --csv \
   --log-file gpurun_out/${TAG}_bench_launches26.csv python bench.py --steps 2 --warmup 3 --no-cpu > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_pr_cb|k_pr_sell|k_pr_finish" -s 30 -c 3 \
   -f -o gpurun_out/${TAG}_prof26 python tools/pr_knobs.py --scale 26 --configs "TAU=1.5" --reps 1 > gpurun_out/${TAG}_prof26.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_pr_cb|k_pr_sell|k_pr_finish" -s 30 -c 3 \
   -f -o gpurun_out/${TAG}_prof22 python tools/pr_knobs.py --scale 22 --configs "TAU=1.5" --reps 1 > gpurun_out/${TAG}_prof22.log 2>&1
ls -la gpurun_out/${TAG}_* | awk '{print $5, $9}'
