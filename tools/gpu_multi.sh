#!/bin/bash
# N-GPU validation: the multi-process tests and the sharded bench line
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TAG="${1:-r2m}"; N="${2:-2}"; SCALE="${3:-26}"; EXTRA="${4:-}"
if [ "$N" = "2" ]; then
  timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/${TAG}_pytest_multi.log 2>&1
  echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_multi.log
  tail -15 gpurun_out/${TAG}_pytest_multi.log
fi
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus $N --steps 5 --warmup 3 --scale $SCALE --diag $EXTRA > gpurun_out/${TAG}_bench_n${N}.json 2> gpurun_out/${TAG}_bench_n${N}.err
echo "bench exit $?"
cut -c1-2500 gpurun_out/${TAG}_bench_n${N}.json
grep -v "^W0\|^\*\*\*\|OMP_NUM" gpurun_out/${TAG}_bench_n${N}.err | tail -12
