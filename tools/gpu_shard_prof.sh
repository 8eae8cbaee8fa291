#!/bin/bash
# ncu --set full of the three sweep kernels of ONE eighth-shard (rank 0 of 8, no peers) on one GPU
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TAG="${1:-shardprof}"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_pr_cb|k_pr_sell|k_pr_finish" -s 75 -c 3 \
   -f -o gpurun_out/${TAG}_w8 python tools/shard_trace.py --scale 26 --world 8 > gpurun_out/${TAG}_w8.log 2>&1
tail -2 gpurun_out/${TAG}_w8.log
ls -la gpurun_out/${TAG}_w8.ncu-rep
