#!/bin/bash
# 2-GPU check: multi-GPU tests + one bench line
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TAG="${1:-r2z}"; N="${2:-2}"
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/${TAG}_pytest_multi.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_multi.log; tail -3 gpurun_out/${TAG}_pytest_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus $N --steps 4 --warmup 3 --diag > gpurun_out/${TAG}_bench_n${N}.json 2> gpurun_out/${TAG}_bench_n${N}.err
echo "bench exit $?"; python -c "
import json; d=json.load(open('gpurun_out/${TAG}_bench_n${N}.json')); print(d['n_gpus'], d['value'], d['ms_per_step'], d['verified'], d['verification']['max_rel_err_vs_single_gpu'], d['config']['multicast'], d['e2e']['value'], d['e2e']['ms_per_step'])"
grep "diag per rank" gpurun_out/${TAG}_bench_n${N}.err | cut -c1-400
