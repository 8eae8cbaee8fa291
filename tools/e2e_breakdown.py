#!/usr/bin/env python3
"""Where does an e2e PageRank step spend its time?  (host CSR -> twin -> plan -> 20 sweeps -> host)"""
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
del g; torch.cuda.empty_cache()
_, hs = bench.pinned_empty(n, np.float32)
cfg = _capi.PageRankConfig(20, 0.0, 0.85, _capi.PR_JACOBI)
it, err = C.c_uint64(0), C.c_double(0)
P = lambda a: a.ctypes.data_as(C.c_void_p)
for rep in range(3):
    h = C.c_void_p(); t0 = time.perf_counter()
    check(lib.gb_digraph_from_csr_u32(0, n, P(oo), P(ot), None, P(io), P(it_), C.byref(h))); t1 = time.perf_counter()
    cfg1 = _capi.PageRankConfig(1, 0.0, 0.85, _capi.PR_JACOBI)
    check(lib.gb_page_rank(h, C.byref(cfg1), P(hs), C.byref(it), C.byref(err))); t2 = time.perf_counter()  # plan + 1 sweep
    check(lib.gb_page_rank(h, C.byref(cfg), P(hs), C.byref(it), C.byref(err))); t3 = time.perf_counter()
    check(lib.gb_graph_free(h)); t4 = time.perf_counter()
    print(f"from_csr {1e3*(t1-t0):.1f} ms ({(8*m+8*n)/(t1-t0)/1e9:.1f} GB/s)  plan+1sweep {1e3*(t2-t1):.1f} ms  20 sweeps+D2H {1e3*(t3-t2):.1f} ms  free {1e3*(t4-t3):.1f} ms")
