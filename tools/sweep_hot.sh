# hot-head size sweep (entries of out_scores mirrored in shared memory) -> GTEPS, ms per sweep
for hot in 0 16384 24576 28672 32768 36864 40960 45056; do
for s in 22 26; do
echo -n "hot=$hot s=$s : "; GB_PR_HOT=$hot python bench.py --scale $s --steps 2 --no-cpu 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['roofline']['mean_launch_ms'],3))"
done; done
