#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python tools/plan_time.py --scale 26 --reps 4; python tools/plan_time.py --scale 22 --reps 4
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r2n_bench26.json 2>gpurun_out/r2n_bench26.err
python -c "
import json; d=json.load(open('gpurun_out/r2n_bench26.json')); print(d['value'], d['verified'], d['roofline']['frac'], d['e2e'])"
tail -2 gpurun_out/r2n_bench26.err
for a in wcc tc sssp; do timeout 400 python bench.py --algo $a --steps 3 --warmup 1 > gpurun_out/r2n_algo_$a.json 2> gpurun_out/r2n_algo_$a.err; cut -c1-900 gpurun_out/r2n_algo_$a.json; tail -2 gpurun_out/r2n_algo_$a.err; done
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "sssp or wcc" 2>&1 | tail -3
