"""CPU oracle for the neo4j-labs/graph hot path — TEST INFRASTRUCTURE ONLY.

ctypes front-end of ``oracle/oracle.c`` (a plain-C restatement of the reference's algorithms,
each function citing the reference file:line it follows).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import this package; the product package ``graph_b200`` never does.

Parity status: the Rust reference cannot be compiled in this image (no cargo/rustc, crates not
vendored), so there is no ``oracle/_ref``.  The restatement is pinned against every golden vector
the reference's own tests hold for this path (``tests/golden/reference_goldens.json`` — see
``tests/test_oracle_goldens.py``).  Large-n PageRank is not pinned by the reference at all (its
multi-threaded sweep is schedule dependent; SURVEY.md §7 hard part 1).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_SRC = _HERE / "oracle.c"
_LIB = _HERE / "liboracle.so"

OUTGOING, INCOMING, UNDIRECTED = 0, 1, 2
UNSORTED, SORTED, DEDUPLICATED = 0, 1, 2


def build(force: bool = False) -> Path:
    """Compile oracle.c -> liboracle.so (gcc, -ffp-contract=off so no FMA sneaks into the f32 sweep)."""
    if force or not _LIB.exists() or _LIB.stat().st_mtime < _SRC.stat().st_mtime:
        tmp = _LIB.with_suffix(f".{os.getpid()}.tmp")
        cmd = ["gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-pthread", "-ffp-contract=off",
               "-fvisibility=hidden", "-Wall", "-Wextra", "-o", str(tmp), str(_SRC), "-lm"]
        subprocess.run(cmd, check=True)
        os.replace(tmp, _LIB)
    return _LIB


_lib = None
_u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(str(build()))
        L = _lib
        L.orc_hardware_threads.restype = C.c_int
        L.orc_graph500_decode.restype = C.c_int64
        L.orc_graph500_decode.argtypes = [_u8p, C.c_uint64, _u32p, _u32p]
        L.orc_edgelist_parse.restype = C.c_int64
        L.orc_edgelist_parse.argtypes = [C.c_char_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_node_count.restype = C.c_uint32
        L.orc_node_count.argtypes = [_u32p, _u32p, C.c_uint64]
        L.orc_csr_build.restype = C.c_uint64
        L.orc_csr_build.argtypes = [_u32p, _u32p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_int, C.c_int,
                                    _u32p, _u32p, C.c_void_p]
        L.orc_make_degree_ordered.restype = None
        L.orc_make_degree_ordered.argtypes = [_u32p, _u32p, C.c_uint32, _u32p, _u32p, C.c_void_p]
        L.orc_in_degree_partition.restype = C.c_uint32
        L.orc_in_degree_partition.argtypes = [_u32p, C.c_uint32, C.c_uint64, C.c_uint32, _u32p]
        pr_common = [_u32p, _u32p, _u32p, C.c_uint32, C.c_uint64, C.c_double, C.c_float]
        pr_tail = [_f32p, C.POINTER(C.c_uint64), C.POINTER(C.c_double)]
        L.orc_page_rank_seq.restype = None
        L.orc_page_rank_seq.argtypes = pr_common + pr_tail
        L.orc_page_rank_jacobi.restype = None
        L.orc_page_rank_jacobi.argtypes = pr_common + [C.c_int] + pr_tail
        L.orc_page_rank_mt.restype = None
        L.orc_page_rank_mt.argtypes = pr_common + [C.c_int] + pr_tail
        L.orc_wcc_afforest.restype = None
        L.orc_wcc_afforest.argtypes = [_u32p, _u32p, _u32p, _u32p, C.c_uint32, C.c_uint64, C.c_uint64,
                                       C.c_uint64, C.c_uint64, C.c_int, _u32p]
        L.orc_wcc_min_label.restype = None
        L.orc_wcc_min_label.argtypes = [_u32p, _u32p, C.c_uint32, _u32p]
        L.orc_dss_ops.restype = None
        L.orc_dss_ops.argtypes = [C.c_uint32, _u32p, C.c_uint64, _u32p, _u32p]
        L.orc_wcc_dss.restype = None
        L.orc_wcc_dss.argtypes = [_u32p, _u32p, _u32p, _u32p, C.c_uint32, C.c_int, C.c_uint64, C.c_uint64,
                                  C.c_uint64, _u32p, _u32p]
        L.orc_sssp_delta_stepping.restype = C.c_int
        L.orc_sssp_delta_stepping.argtypes = [_u32p, _u32p, _f32p, C.c_uint32, C.c_uint64, C.c_float, _f32p]
        L.orc_sssp_bellman_ford.restype = C.c_int
        L.orc_sssp_bellman_ford.argtypes = [_u32p, _u32p, _f32p, C.c_uint32, C.c_uint64, _f32p]
        L.orc_triangle_count.restype = C.c_uint64
        L.orc_triangle_count.argtypes = [_u32p, _u32p, C.c_uint32, C.c_int]
        L.orc_rmat_edges.restype = None
        L.orc_rmat_edges.argtypes = [C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint64, _u32p, _u32p]
        L.orc_rmat_weights.restype = None
        L.orc_rmat_weights.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, _f32p]
        L.orc_now_seconds.restype = C.c_double
    return _lib


def _u32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.uint32)


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def _opt(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def hardware_threads() -> int:
    return int(lib().orc_hardware_threads())


# ---- inputs ---------------------------------------------------------------------------------
def graph500_decode(raw: bytes):
    """(src, dst, node_count) of a packed Graph500 file (input/graph500.rs:63-127)."""
    buf = np.frombuffer(raw, dtype=np.uint8)
    m = len(raw) // 12
    src = np.empty(m, np.uint32)
    dst = np.empty(m, np.uint32)
    got = lib().orc_graph500_decode(np.ascontiguousarray(buf), len(raw), src, dst)
    if got < 0:
        raise ValueError("Graph500 id does not fit u32")
    return src, dst, m // 16


def edgelist_parse(text: bytes, with_values: bool = False):
    """(src, dst[, values]) of a text edge list (input/edgelist.rs:181-279)."""
    m = lib().orc_edgelist_parse(text, len(text), None, None, None)
    src = np.empty(m, np.uint32)
    dst = np.empty(m, np.uint32)
    val = np.empty(m, np.float32) if with_values else None
    lib().orc_edgelist_parse(text, len(text), src.ctypes.data_as(C.c_void_p),
                             dst.ctypes.data_as(C.c_void_p), _opt(val))
    return (src, dst, val) if with_values else (src, dst)


def node_count(src, dst) -> int:
    src, dst = _u32(src), _u32(dst)
    return int(lib().orc_node_count(src, dst, len(src)))


# ---- CSR ------------------------------------------------------------------------------------
def csr_build(src, dst, n: int, direction: int, layout: int, weights=None):
    """offsets[n+1], targets (and weights) following csr.rs:124-221 on one thread."""
    src, dst = _u32(src), _u32(dst)
    m = len(src)
    w = None if weights is None else _f32(weights)
    cap = (2 * m if direction == UNDIRECTED else m) or 1
    off = np.zeros(n + 1, np.uint32)
    tgt = np.empty(cap, np.uint32)
    wout = np.empty(cap, np.float32) if w is not None else None
    ln = lib().orc_csr_build(src, dst, _opt(w), m, n, direction, layout, off, tgt, _opt(wout))
    tgt = tgt[:ln].copy()
    if w is not None:
        return off, tgt, wout[:ln].copy()
    return off, tgt


def make_degree_ordered(off, tgt):
    """(new_offsets, new_targets, new_id) — graph_ops.rs:511-638."""
    off, tgt = _u32(off), _u32(tgt)
    n = len(off) - 1
    noff = np.zeros(n + 1, np.uint32)
    ntgt = np.empty(max(len(tgt), 1), np.uint32)
    nid = np.empty(max(n, 1), np.uint32)
    tg = tgt if len(tgt) else np.zeros(1, np.uint32)
    lib().orc_make_degree_ordered(off, tg, n, noff, ntgt, nid.ctypes.data_as(C.c_void_p))
    return noff, ntgt[:len(tgt)].copy(), nid[:n].copy()


def in_degree_partition(in_off, parts: int):
    in_off = _u32(in_off)
    n = len(in_off) - 1
    ranges = np.zeros(parts + 1, np.uint32)
    cnt = lib().orc_in_degree_partition(in_off, n, int(in_off[n]), parts, ranges)
    return ranges[:cnt + 1].copy()


# ---- algorithms -----------------------------------------------------------------------------
def _pr(fn, in_off, in_tgt, out_off, max_iterations, tolerance, damping, extra):
    in_off, in_tgt, out_off = _u32(in_off), _u32(in_tgt), _u32(out_off)
    n = len(in_off) - 1
    if len(in_tgt) == 0:
        in_tgt = np.zeros(1, np.uint32)
    scores = np.empty(max(n, 1), np.float32)
    it = C.c_uint64(0)
    err = C.c_double(0.0)
    fn(in_off, in_tgt, out_off, n, max_iterations, tolerance, damping, *extra, scores,
       C.byref(it), C.byref(err))
    return scores[:n], int(it.value), float(err.value)


def page_rank_seq(in_off, in_tgt, out_off, max_iterations=20, tolerance=1e-4, damping=0.85):
    """The reference's sweep as one thread runs it (page_rank.rs:58-168) — bit-exact goldens."""
    return _pr(lib().orc_page_rank_seq, in_off, in_tgt, out_off, max_iterations, tolerance, damping, [])


def page_rank_jacobi(in_off, in_tgt, out_off, max_iterations=20, tolerance=1e-4, damping=0.85,
                     acc64=True):
    """Same update rule, double-buffered schedule; acc64 -> f64 row sums rounded once."""
    return _pr(lib().orc_page_rank_jacobi, in_off, in_tgt, out_off, max_iterations, tolerance,
               damping, [1 if acc64 else 0])


def page_rank_mt(in_off, in_tgt, out_off, max_iterations=20, tolerance=1e-4, damping=0.85, threads=0):
    """The reference's multi-threaded in-place sweep (chunk 16384, atomic claim) — timed baseline."""
    return _pr(lib().orc_page_rank_mt, in_off, in_tgt, out_off, max_iterations, tolerance, damping,
               [threads])


def wcc_afforest(out_off, out_tgt, in_off, in_tgt, chunk_size=16384, neighbor_rounds=2,
                 sampling_size=1024, rng_seed=42, threads=1):
    out_off, out_tgt, in_off, in_tgt = _u32(out_off), _u32(out_tgt), _u32(in_off), _u32(in_tgt)
    n = len(out_off) - 1
    if len(out_tgt) == 0:
        out_tgt = np.zeros(1, np.uint32)
        in_tgt = np.zeros(1, np.uint32)
    comp = np.empty(max(n, 1), np.uint32)
    lib().orc_wcc_afforest(out_off, out_tgt, in_off, in_tgt, n, chunk_size, neighbor_rounds,
                           sampling_size, rng_seed, threads, comp)
    return comp[:n]


def dss_ops(n: int, pairs):
    """DisjointSetStruct (dss.rs:38-116): unions in order -> (to_vec parents, find(i) for every i)."""
    pairs = _u32(np.asarray(pairs, np.uint32).reshape(-1))
    if len(pairs) == 0:
        pairs = np.zeros(2, np.uint32)
        npairs = 0
    else:
        npairs = len(pairs) // 2
    parents, finds = np.empty(max(n, 1), np.uint32), np.empty(max(n, 1), np.uint32)
    lib().orc_dss_ops(n, pairs, npairs, parents, finds)
    return parents[:n], finds[:n]


def wcc_dss(out_off, out_tgt, in_off, in_tgt, variant="afforest_dss", neighbor_rounds=2, sampling_size=1024,
            rng_seed=42):
    """wcc_baseline / wcc_afforest_dss (wcc.rs:103-156) on one thread -> (to_vec(), component(i) for every i)."""
    out_off, out_tgt, in_off, in_tgt = _u32(out_off), _u32(out_tgt), _u32(in_off), _u32(in_tgt)
    n = len(out_off) - 1
    if len(out_tgt) == 0:
        out_tgt = np.zeros(1, np.uint32)
    if len(in_tgt) == 0:
        in_tgt = np.zeros(1, np.uint32)
    to_vec, comp = np.empty(max(n, 1), np.uint32), np.empty(max(n, 1), np.uint32)
    lib().orc_wcc_dss(out_off, out_tgt, in_off, in_tgt, n, {"baseline": 0, "afforest_dss": 1}[variant],
                      neighbor_rounds, sampling_size, rng_seed, to_vec, comp)
    return to_vec[:n], comp[:n]


def wcc_min_label(out_off, out_tgt):
    out_off, out_tgt = _u32(out_off), _u32(out_tgt)
    n = len(out_off) - 1
    if len(out_tgt) == 0:
        out_tgt = np.zeros(1, np.uint32)
    comp = np.empty(max(n, 1), np.uint32)
    lib().orc_wcc_min_label(out_off, out_tgt, n, comp)
    return comp[:n]


def sssp_delta_stepping(off, tgt, w, start: int, delta: float):
    off, tgt, w = _u32(off), _u32(tgt), _f32(w)
    n = len(off) - 1
    if len(tgt) == 0:
        tgt, w = np.zeros(1, np.uint32), np.zeros(1, np.float32)
    dist = np.empty(max(n, 1), np.float32)
    if lib().orc_sssp_delta_stepping(off, tgt, w, n, start, delta, dist) != 0:
        raise IndexError("start node out of range")
    return dist[:n]


def sssp_bellman_ford(off, tgt, w, start: int):
    off, tgt, w = _u32(off), _u32(tgt), _f32(w)
    n = len(off) - 1
    if len(tgt) == 0:
        tgt, w = np.zeros(1, np.uint32), np.zeros(1, np.float32)
    dist = np.empty(max(n, 1), np.float32)
    if lib().orc_sssp_bellman_ford(off, tgt, w, n, start, dist) != 0:
        raise IndexError("start node out of range")
    return dist[:n]


def triangle_count(off, tgt, threads=1) -> int:
    off, tgt = _u32(off), _u32(tgt)
    if len(tgt) == 0:
        tgt = np.zeros(1, np.uint32)
    return int(lib().orc_triangle_count(off, tgt, len(off) - 1, threads))


# ---- synthetic workload ---------------------------------------------------------------------
def rmat_edges(scale: int, seed: int = 42, first: int = 0, count: int | None = None, edge_factor=16):
    if count is None:
        count = edge_factor << scale
    src = np.empty(max(count, 1), np.uint32)
    dst = np.empty(max(count, 1), np.uint32)
    lib().orc_rmat_edges(scale, seed, first, count, src, dst)
    return src[:count], dst[:count]


def rmat_weights(seed: int, first: int, count: int):
    w = np.empty(max(count, 1), np.float32)
    lib().orc_rmat_weights(seed, first, count, w)
    return w[:count]


def now() -> float:
    return float(lib().orc_now_seconds())
