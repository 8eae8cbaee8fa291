#!/bin/bash
# one development iteration on the GPU: PageRank parity subset, knob timing, per-kernel times
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TAG="${1:-r2h}"
CF26="${2:-B=32768,TAU=2;B=32768,TAU=2,DUAL=0;B=49152,TAU=2,DUAL=0;B=49152,TAU=2;B=32768,TAU=1.5}"
CF22="${3:-B=32768,TAU=2;B=32768,TAU=2,DUAL=0;B=49152,TAU=2,DUAL=0;B=16384,TAU=2}"
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "page_rank or shard" > gpurun_out/${TAG}_pytest_pr.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_pr.log
tail -4 gpurun_out/${TAG}_pytest_pr.log
timeout 300 python tools/pr_knobs.py --scale 22 --configs "$CF22" > gpurun_out/${TAG}_knobs22.jsonl 2> gpurun_out/${TAG}_knobs22.err
cut -c1-110 gpurun_out/${TAG}_knobs22.jsonl
timeout 600 python tools/pr_knobs.py --scale 26 --configs "$CF26" > gpurun_out/${TAG}_knobs26.jsonl 2> gpurun_out/${TAG}_knobs26.err
cut -c1-110 gpurun_out/${TAG}_knobs26.jsonl
tail -3 gpurun_out/${TAG}_knobs26.err
bash tools/gpu_launches.sh ${TAG} "B=32768,TAU=2,DUAL=0" "26 22" 2>&1 | grep -v "^{" 
