#!/bin/bash
# full single-GPU validation: every -m gpu test, smoke, the default bench line, the reference arm
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TAG="${1:-r2k}"
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -6 gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as e; e.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -2 gpurun_out/${TAG}_smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench26.json 2> gpurun_out/${TAG}_bench26.err
cut -c1-1500 gpurun_out/${TAG}_bench26.json; tail -3 gpurun_out/${TAG}_bench26.err
timeout 600 python bench.py --steps 5 --warmup 3 --scale 22 > gpurun_out/${TAG}_bench22.json 2> gpurun_out/${TAG}_bench22.err
cut -c1-600 gpurun_out/${TAG}_bench22.json; tail -3 gpurun_out/${TAG}_bench22.err
