"""bench_wcc_multi.py — multi-GPU WCC (ShardedWcc) next to the single-GPU gb_wcc on the same graph.
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_wcc_multi.py --scale 24
One JSON line from rank 0: device time (max over ranks, CUDA events), bit-equality with the 1-GPU labels."""
import argparse, json, os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=int, default=24)
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    import graph_b200 as gb
    from graph_b200.multigpu import ShardedWcc
    gb.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    g = gb.DiGraph.rmat(a.scale, 16, 42, gb.Layout.Sorted)
    n, m = g.node_count(), g.edge_count()
    sw = ShardedWcc(g)
    sw.run()
    times = []
    for _ in range(a.reps):
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        comp = sw.run()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        times.append(float(t.item()))
    single_ms, equal = None, None
    if rank == 0:
        g.wcc()
        res = g.wcc()
        single_ms = g.last_timing()["total_ms"]
        equal = bool((comp.cpu().numpy().view(np.uint32) == res.components()).all())
        ms = float(np.median(times))
        print(json.dumps({"algo": "wcc_afforest sharded", "scale": a.scale, "n_gpus": world, "ms": ms,
                          "g_edges_per_s": m / (ms * 1e-3) / 1e9, "ms_all": [round(t, 3) for t in times],
                          "single_gpu_ms": single_ms, "bit_equal_to_single_gpu": equal,
                          "exchange": "2 NCCL all-gathers of parent[n] (4n bytes per rank each) + P-1 forest merges each"}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
