"""check_scores_host.py — one-rank sanity of ShardedPageRank.scores_host(reuse=True) on a single GPU."""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29577")
import numpy as np, torch, torch.distributed as dist
torch.cuda.set_device(0)
dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
import graph_b200 as gb
from graph_b200.multigpu import ShardedPageRank
g = gb.DiGraph.rmat(16, seed=42, layout=gb.Layout.Sorted)
spr = ShardedPageRank(g, exchange="allgather")
spr.run(20, 0.85, 0.0)
a = spr.scores_host().copy()
b = spr.scores_host(reuse=True).copy()
c = spr.scores_host(reuse=True)
single = g.page_rank(max_iterations=20, tolerance=0.0, mode="jacobi").scores()
print("scores_host ok" if (a == b).all() and (a == c).all() and np.max(np.abs(a - single) / single) <= 1e-6 else "MISMATCH",
      c.flags["C_CONTIGUOUS"], spr._host_scores.is_pinned())
dist.destroy_process_group()
