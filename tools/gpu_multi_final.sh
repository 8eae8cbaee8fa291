#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TAG="${1:-r2y}"; N="${2:-8}"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus $N --steps 5 --warmup 3 --diag > gpurun_out/${TAG}_bench_n${N}.json 2> gpurun_out/${TAG}_bench_n${N}.err
echo "bench exit $?"; python -c "
import json; d=json.load(open('gpurun_out/${TAG}_bench_n${N}.json')); print(d['n_gpus'], d['value'], d['ms_per_step'], d['verified'], d['verification']['max_rel_err_vs_single_gpu'], d['config']['multicast'], d['e2e']['value'], d['e2e']['ms_per_step'])"
grep "diag per rank" gpurun_out/${TAG}_bench_n${N}.err | cut -c1-700
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tools/bench_wcc_multi.py --scale 24 2>/dev/null | grep "^{" > gpurun_out/${TAG}_wcc_n${N}.json; cat gpurun_out/${TAG}_wcc_n${N}.json
timeout 600 python tools/bench_comm.py --scale 26 --gpus $N 2>&1 | grep "^{" > gpurun_out/${TAG}_comm_n${N}.json; cat gpurun_out/${TAG}_comm_n${N}.json
