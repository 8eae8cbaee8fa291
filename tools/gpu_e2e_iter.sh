#!/bin/bash
# one iteration on the end-to-end path: PageRank parity subset, layout build time, e2e bench with feed knobs
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TAG="${1:-e2e}"
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "page_rank or shard or communicator or column" > gpurun_out/${TAG}_pytest_pr.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_pr.log
tail -5 gpurun_out/${TAG}_pytest_pr.log
timeout 300 python tools/plan_time.py --scale 26 > gpurun_out/${TAG}_plan26.log 2>&1
tail -4 gpurun_out/${TAG}_plan26.log
for K in ${2:-8 16 0}; do
  GB_PR_FEED_CHUNKS=$K timeout 400 python bench.py --no-cpu --steps 4 --warmup 3 > gpurun_out/${TAG}_bench_k$K.json 2> gpurun_out/${TAG}_bench_k$K.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_bench_k$K.json").read().strip().splitlines()[-1])
    print("K=$K value", d["value"], "frac", d["roofline"]["frac"], "e2e", d["e2e"], "verified", d.get("config", {}).get("verified"))
except Exception as ex:
    print("K=$K failed", ex)
PY
  tail -2 gpurun_out/${TAG}_bench_k$K.err
done
