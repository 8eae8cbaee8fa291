"""bench_comm.py — gb_page_rank_multi (single process, all GPUs of the box, no torch.distributed / NCCL).
  python tools/bench_comm.py --scale 26 --gpus 8
Wall time of the whole call: shard layouts are built inside it (nothing cached), then 20 sweeps."""
import argparse, json, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
ap = argparse.ArgumentParser(); ap.add_argument("--scale", type=int, default=26); ap.add_argument("--gpus", type=int, default=2)
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
import numpy as np, graph_b200 as gb
graphs = []
for d in range(a.gpus):
    gb.set_device(d)
    graphs.append(gb.DiGraph.rmat(a.scale, 16, 42, gb.Layout.Sorted))
gb.set_device(0)
comm = gb.Comm(list(range(a.gpus)))
m = graphs[0].edge_count()
comm.page_rank(graphs, max_iterations=20, tolerance=0.0)
ts = []
for _ in range(a.reps):
    t0 = time.perf_counter(); pr = comm.page_rank(graphs, max_iterations=20, tolerance=0.0); ts.append(time.perf_counter() - t0)
single = graphs[0].page_rank(max_iterations=20, tolerance=0.0, mode="jacobi").scores()
print(json.dumps({"what": "gb_page_rank_multi, single process", "scale": a.scale, "n_gpus": a.gpus, "multicast": comm.multicast,
                  "call_ms": [round(t * 1e3, 2) for t in ts], "gteps_whole_call": m * 20 / min(ts) / 1e9,
                  "max_rel_err_vs_single_gpu": float(np.max(np.abs(pr.scores() - single) / single))}))
