#!/bin/bash
# first GPU trip of round 2: PageRank parity, a sanitizer pass over the new kernels, knob timing
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TAG="${1:-r2a}"
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "page_rank or shard" > gpurun_out/${TAG}_pytest_pr.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_pr.log
tail -15 gpurun_out/${TAG}_pytest_pr.log
cat > /tmp/san.py <<'PY'
import os, numpy as np
os.environ.update(GB_PR_BLOCK="1024", GB_PR_CHUNK="32", GB_PR_TAU="1", GB_PR_MIN_BLOCK="0")
import oracle, graph_b200 as gb
src, dst = oracle.rmat_edges(13, seed=3)
g = gb.DiGraph.from_numpy(np.stack([src, dst], 1), layout=gb.Layout.Sorted)
pr = g.page_rank(max_iterations=3, tolerance=0.0, mode="jacobi")
print("sanitizer run ok", pr.error, g.page_rank_plan_info())
PY
PYTHONPATH=$PWD timeout 600 compute-sanitizer --tool memcheck python /tmp/san.py > gpurun_out/${TAG}_sanitizer.log 2>&1
echo "sanitizer exit $?" >> gpurun_out/${TAG}_sanitizer.log
tail -5 gpurun_out/${TAG}_sanitizer.log
timeout 300 python tools/pr_knobs.py --scale 22 --configs "B=32768,TAU=2;B=32768,TAU=2,DUAL=0;B=49152,TAU=2,DUAL=0;B=24576,TAU=2;B=32768,TAU=1.5" > gpurun_out/${TAG}_knobs22.jsonl 2> gpurun_out/${TAG}_knobs22.err
cat gpurun_out/${TAG}_knobs22.jsonl | cut -c1-260
timeout 600 python tools/pr_knobs.py --scale 26 --configs "B=32768,TAU=2;B=32768,TAU=2,DUAL=0;B=49152,TAU=2,DUAL=0;B=49152,TAU=3,DUAL=0;B=57344,TAU=2,DUAL=0;B=32768,TAU=1.5;B=32768,TAU=2,MINB=0" > gpurun_out/${TAG}_knobs26.jsonl 2> gpurun_out/${TAG}_knobs26.err
cat gpurun_out/${TAG}_knobs26.jsonl | cut -c1-260
tail -3 gpurun_out/${TAG}_knobs26.err
