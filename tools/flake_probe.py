"""flake_probe.py — repeat the JACOBI run on one resident graph and report runs whose bits differ from the first."""
import argparse, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
ap = argparse.ArgumentParser(); ap.add_argument("--scales", default="16,18,20"); ap.add_argument("--reps", type=int, default=60)
a = ap.parse_args()
import os
import graph_b200 as gb
graphs = {sc: gb.DiGraph.rmat(sc, 16, 7, gb.Layout.Sorted) for sc in [int(x) for x in a.scales.split(",")]}
for dbg, sc in [(d, sc) for d in os.environ.get("PROBE_DEBUG", "0,1,2,4,7").split(",") for sc in graphs]:
    os.environ["GB_PR_DEBUG"] = dbg
    g = graphs[sc]
    info = g.page_rank_plan_info()
    first = g.page_rank(max_iterations=20, tolerance=0.0, mode="jacobi").scores().copy()
    indeg = None
    bad = 0
    for r in range(a.reps):
        s = g.page_rank(max_iterations=20, tolerance=0.0, mode="jacobi").scores()
        d = np.nonzero(s.view(np.uint32) != first.view(np.uint32))[0]
        if len(d):
            bad += 1
            if bad <= 3:
                if indeg is None:
                    indeg = np.array([g.in_degree(int(v)) for v in d[:8]])
                rel = np.abs(s[d] - first[d]) / first[d]
                print(f"  debug {dbg} scale {sc} rep {r}: {len(d)} vertices differ, first ids {d[:8].tolist()} in-degrees {[g.in_degree(int(v)) for v in d[:8]]} max rel {rel.max():.2e}")
    print(f"debug {dbg} scale {sc}: {bad}/{a.reps} runs differ; hot_blocks {info.get('hot_blocks')}")
