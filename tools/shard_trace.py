"""shard_trace.py — per-kernel times of one shard's sweeps on ONE GPU (no peers): what rank 0 of a
`world`-way run spends in its kernels, without needing `world` GPUs.
  GB_PR_TRACE=1 python tools/shard_trace.py --scale 26 --world 8"""
import argparse, os, sys, time
from pathlib import Path
os.environ["GB_PR_TRACE"] = "1"
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
ap = argparse.ArgumentParser(); ap.add_argument("--scale", type=int, default=26); ap.add_argument("--world", type=int, default=8)
ap.add_argument("--sweeps", type=int, default=20)
a = ap.parse_args()
import torch, graph_b200 as gb
from graph_b200.multigpu import CudaShardBackend
g = gb.DiGraph.rmat(a.scale, 16, 42, gb.Layout.Sorted)
n = g.node_count()
b = CudaShardBackend(g, 0, a.world)
x = [torch.zeros(n, dtype=torch.float32, device="cuda") for _ in range(2)]
scores = torch.zeros(n, dtype=torch.float32, device="cuda"); err = torch.zeros(1, dtype=torch.float64, device="cuda")
for rep in range(2):
    b.init(0.85, x[0], x[1], scores)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for sw in range(1, a.sweeps + 1):
        b.step(0.85, sw, x[(sw - 1) & 1], x[sw & 1], None, scores, err)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
print({"scale": a.scale, "world": a.world, "ms_per_sweep_wall": round(dt / a.sweeps * 1e3, 4), "knobs": {k: v for k, v in os.environ.items() if k.startswith("GB_PR_") and k != "GB_PR_TRACE"}, "tasks": b.stats["tasks"], "chunk_groups": b.stats["chunk_groups"], "hot_blocks": b.stats["hot_blocks"]})
del b   # prints the trace
