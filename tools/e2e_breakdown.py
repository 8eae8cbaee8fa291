#!/usr/bin/env python3
"""Per-step wall time of the e2e path (gb_page_rank_csr_u32 on pinned host CSR) and of its parts."""
import ctypes as C, sys, time
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import graph_b200 as gb
from graph_b200 import _capi
from graph_b200._capi import lib, check
import bench
scale = int(sys.argv[1]) if len(sys.argv) > 1 else 26
n = 1 << scale; m = 16 * n
g = gb.DiGraph.rmat(scale)
(oo, ot, io, it_), keep = bench.host_csr_from_device(g)
print("pinned:", [k.is_pinned() for k in keep])
del g; torch.cuda.empty_cache()
_, hs = bench.pinned_empty(n, np.float32)
cfg = _capi.PageRankConfig(20, 0.0, 0.85, _capi.PR_JACOBI)
it, err = C.c_uint64(0), C.c_double(0)
P = lambda a: a.ctypes.data_as(C.c_void_p)
for rep in range(4):
    t0 = time.perf_counter()
    check(lib.gb_page_rank_csr_u32(0, n, P(io), P(it_), P(oo), C.byref(cfg), P(hs), C.byref(it), C.byref(err)))
    t1 = time.perf_counter()
    print(f"one-shot step {1e3*(t1-t0):.1f} ms -> {m*20/(t1-t0)/1e9:.1f} GTEPS")
d = torch.empty(m, dtype=torch.int32, device="cuda")
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    d.copy_(torch.from_numpy(it_.view(np.int32)), non_blocking=True); torch.cuda.synchronize()
    t1 = time.perf_counter(); print(f"H2D 4.3 GB: {1e3*(t1-t0):.1f} ms = {4*m/(t1-t0)/1e9:.1f} GB/s")
