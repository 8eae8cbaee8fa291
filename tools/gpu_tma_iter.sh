#!/bin/bash
# iteration on the sweep kernels: parity subset, per-kernel trace at scale 22 / 26 / one eighth-shard, bench line
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TAG="${1:-tma}"
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "page_rank or shard or communicator or column" > gpurun_out/${TAG}_pytest_pr.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_pr.log
tail -5 gpurun_out/${TAG}_pytest_pr.log
for S in 22 26; do
  GB_PR_TRACE=1 timeout 300 python tools/pr_knobs.py --scale $S --configs "B=49152,TAU=1.5" > gpurun_out/${TAG}_knobs$S.jsonl 2> gpurun_out/${TAG}_knobs$S.err
  cut -c1-200 gpurun_out/${TAG}_knobs$S.jsonl; grep "gb trace" gpurun_out/${TAG}_knobs$S.err | tail -2
done
timeout 300 python tools/shard_trace.py --scale 26 --world 8 > gpurun_out/${TAG}_shard8.log 2>&1
tail -3 gpurun_out/${TAG}_shard8.log
true
true
