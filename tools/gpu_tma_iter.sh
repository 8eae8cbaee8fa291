#!/bin/bash
# iteration on the sweep kernels: GPU tests, per-kernel trace at scale 22 / 26 / one eighth-shard under knob sets
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TAG="${1:-tma}"
CFGS="${2:-B=49152,TAU=1.5}"
SUITE="${3:-subset}"
if [ "$SUITE" = "full" ]; then
  timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/${TAG}_pytest_pr.log 2>&1
else
  timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "page_rank or shard or communicator or column" > gpurun_out/${TAG}_pytest_pr.log 2>&1
fi
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_pr.log
tail -5 gpurun_out/${TAG}_pytest_pr.log
for S in 22 26; do
  GB_PR_TRACE=1 timeout 300 python tools/pr_knobs.py --scale $S --configs "$CFGS" > gpurun_out/${TAG}_knobs$S.jsonl 2> gpurun_out/${TAG}_knobs$S.err
  cut -c1-150 gpurun_out/${TAG}_knobs$S.jsonl; grep "gb trace" gpurun_out/${TAG}_knobs$S.err
done
IFS=';' read -ra ARR <<< "$CFGS"
for CFG in "${ARR[@]}"; do
  T=$(echo "$CFG" | tr ',' '\n' | grep '^TASK=' | cut -d= -f2)
  GB_PR_TASK_CHUNKS=${T:-32} timeout 300 python tools/shard_trace.py --scale 26 --world 8 > gpurun_out/${TAG}_shard8_${T:-32}.log 2>&1
  tail -2 gpurun_out/${TAG}_shard8_${T:-32}.log
done
