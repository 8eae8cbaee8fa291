"""graph_b200 — B200-native drop-in for the CSR hot path of neo4j-labs/graph.

The public names mirror the reference's Python module ``graph_mate`` (crates/mate/graph_mate.pyi:
``DiGraph``, ``Graph``, ``Layout``, ``FileFormat``, ``PageRankResult``, ``WccResult``,
``TriangleCountResult``) so that the reference's own pytest suite reads the same against this
package.  Every algorithm call goes through the C ABI of ``libgraph_b200.so``
(include/graph_b200.h) and runs on the GPU; there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import sys
import time
from pathlib import Path

import numpy as np

from . import _capi
from ._capi import GraphB200Error, check, lib

__all__ = ["DiGraph", "Graph", "Layout", "FileFormat", "PageRankResult", "WccResult",
           "TriangleCountResult", "SsspResult", "PageRankConfig", "WccConfig", "DeltaSteppingConfig",
           "GraphB200Error", "device_count", "set_device", "write_graph500"]

_device = 0


def device_count() -> int:
    return int(lib.gb_device_count())


def set_device(index: int) -> None:
    """CUDA device new graphs are created on (one process per GPU sets this to LOCAL_RANK)."""
    global _device
    _device = int(index)


class Comm:
    """Single-process multi-GPU communicator (gb_comm_*): one host thread, N devices, no torch / NCCL.

        comm = Comm([0, 1])
        graphs = [DiGraph.rmat(22) under set_device(d) for d in comm.devices]   # the same graph on each
        result = comm.page_rank(graphs, max_iterations=20, tolerance=0.0)
    """

    def __init__(self, devices):
        self.devices = [int(d) for d in devices]
        arr = (C.c_int * len(self.devices))(*self.devices)
        self._c = C.c_void_p()
        check(lib.gb_comm_init(len(self.devices), arr, C.byref(self._c)))

    def __del__(self):
        c, self._c = getattr(self, "_c", None), None
        if c:
            try:
                lib.gb_comm_free(c)
            except Exception:
                pass

    @property
    def multicast(self) -> bool:
        n, mc = C.c_int(0), C.c_int(0)
        check(lib.gb_comm_info(self._c, C.byref(n), C.byref(mc)))
        return bool(mc.value)

    def page_rank(self, graphs, *, max_iterations: int = 20, tolerance: float = 1e-4, damping_factor: float = 0.85):
        """page_rank over the communicator's devices (JACOBI schedule); graphs[i] must live on devices[i]."""
        if len(graphs) != len(self.devices):
            raise ValueError("one graph per device of the communicator")
        cfg = _capi.PageRankConfig(int(max_iterations), float(tolerance), float(damping_factor), _capi.PR_JACOBI)
        arr = (C.c_void_p * len(graphs))(*[g._g for g in graphs])
        scores = np.empty(graphs[0].node_count(), np.float32)
        it, err = C.c_uint64(0), C.c_double(0.0)

        def go():
            check(lib.gb_page_rank_multi(self._c, arr, C.byref(cfg), _ptr(scores), C.byref(it), C.byref(err)))
        _, micros = _timed(go)
        return PageRankResult(scores, int(it.value), float(err.value), micros)


# ---- enums (crates/mate/src/graphs/mod.rs Layout / FileFormat; csr.rs:35-45) --------------------
class _Enum:
    def __init__(self, cls_name: str, name: str, value: int):
        self._cls, self.name, self.value = cls_name, name, value

    def __repr__(self):
        return f"{self._cls}.{self.name}"

    def __int__(self):
        return self.value


class Layout:
    """How neighbor lists are organised inside the CSR target array (csr.rs:35-45)."""
    Unsorted = _Enum("Layout", "Unsorted", _capi.LAYOUT_UNSORTED)
    Sorted = _Enum("Layout", "Sorted", _capi.LAYOUT_SORTED)
    Deduplicated = _Enum("Layout", "Deduplicated", _capi.LAYOUT_DEDUPLICATED)


class FileFormat:
    Graph500 = _Enum("FileFormat", "Graph500", 0)
    EdgeList = _Enum("FileFormat", "EdgeList", 1)


def _layout_value(layout) -> int:
    if layout is None:
        return _capi.LAYOUT_UNSORTED  # CsrLayout::default(), csr.rs:35-45
    if isinstance(layout, _Enum) and layout._cls == "Layout":
        return layout.value
    raise TypeError(f"layout must be a graph_b200.Layout, got {layout!r}")


# ---- configs (plain structs with the reference defaults) -----------------------------------------
class PageRankConfig:
    """crates/algos/src/page_rank.rs:14-56"""
    DEFAULT_MAX_ITERATIONS = 20
    DEFAULT_TOLERANCE = 1e-4
    DEFAULT_DAMPING_FACTOR = 0.85

    def __init__(self, max_iterations=DEFAULT_MAX_ITERATIONS, tolerance=DEFAULT_TOLERANCE,
                 damping_factor=DEFAULT_DAMPING_FACTOR):
        self.max_iterations, self.tolerance, self.damping_factor = max_iterations, tolerance, damping_factor


class WccConfig:
    """crates/algos/src/wcc.rs:40-79"""
    DEFAULT_CHUNK_SIZE = 16384
    DEFAULT_NEIGHBOR_ROUNDS = 2
    DEFAULT_SAMPLING_SIZE = 1024

    def __init__(self, chunk_size=DEFAULT_CHUNK_SIZE, neighbor_rounds=DEFAULT_NEIGHBOR_ROUNDS,
                 sampling_size=DEFAULT_SAMPLING_SIZE):
        self.chunk_size, self.neighbor_rounds, self.sampling_size = chunk_size, neighbor_rounds, sampling_size


class DeltaSteppingConfig:
    """crates/algos/src/sssp.rs:18-36"""

    def __init__(self, start_node: int, delta: float):
        self.start_node, self.delta = start_node, delta


# ---- results (crates/mate/src/{page_rank,wcc,triangle_count}.rs) ----------------------------------
def _took(micros: int) -> str:
    return f"{micros / 1000.0:.3f}ms" if micros >= 1000 else f"{micros}µs"


class PageRankResult:
    def __init__(self, scores, ran_iterations, error, micros):
        scores.flags.writeable = False
        self._scores, self.ran_iterations, self.error, self.micros = scores, ran_iterations, error, micros

    def scores(self) -> np.ndarray:
        return self._scores

    def __repr__(self):
        return (f'PageRankResult {{ scores: "[... {len(self._scores)} values]", ran_iterations: '
                f"{self.ran_iterations}, error: {self.error}, took: {_took(self.micros)} }}")


class WccResult:
    def __init__(self, components, micros):
        components.flags.writeable = False
        self._components, self.micros = components, micros

    def components(self) -> np.ndarray:
        return self._components

    def __repr__(self):
        return f'WccResult {{ components: "[... {len(self._components)} values]", took: {_took(self.micros)} }}'


class TriangleCountResult:
    def __init__(self, triangles, micros):
        self.triangles, self.micros = triangles, micros

    def __repr__(self):
        return f"TriangleCountResult {{ triangles: {self.triangles}, took: {_took(self.micros)} }}"


class SsspResult:
    def __init__(self, distances, micros):
        distances.flags.writeable = False
        self._distances, self.micros = distances, micros

    def distances(self) -> np.ndarray:
        return self._distances

    def __repr__(self):
        return f'SsspResult {{ distances: "[... {len(self._distances)} values]", took: {_took(self.micros)} }}'


# ---- input files (crates/builder/src/input/{graph500,edgelist}.rs) --------------------------------
# parsed by the library's native multi-threaded readers (csrc/io.cu)
def _read_graph500(path) -> tuple[np.ndarray, np.ndarray, int]:
    """Packed 12-byte edges {v0_low, v1_low, high} (graph500.rs:111-127); node_count = edges/16 (:74)."""
    raw = np.fromfile(path, dtype=np.uint8)
    m = raw.size // 12
    src = np.empty(m, np.uint32)
    dst = np.empty(m, np.uint32)
    got, n = C.c_uint64(0), C.c_uint32(0)
    check(lib.gb_graph500_decode(_ptr(raw) if raw.size else None, raw.size, _ptr(src), _ptr(dst), C.byref(got),
                                 C.byref(n)))
    return src, dst, int(n.value)


def write_graph500(path, src, dst) -> None:
    """Writes edges as the packed Graph500 file the reference reads (`app -f graph500 --use-32-bit`,
    `DiGraph.load(path)`); the reader derives node_count = edge_count / 16 (graph500.rs:74)."""
    src = np.ascontiguousarray(src, dtype=np.uint32)
    dst = np.ascontiguousarray(dst, dtype=np.uint32)
    if len(src) != len(dst):
        raise ValueError("src and dst must have the same length")
    raw = np.empty(12 * len(src), np.uint8)
    check(lib.gb_graph500_encode(_ptr(src), _ptr(dst), len(src), _ptr(raw)))
    raw.tofile(path)


def _read_edge_list(path, with_values=False):
    """Text lines `<src> <dst>[ <value>]` with \\n or \\r\\n endings (edgelist.rs:181-279)."""
    text = Path(path).read_bytes()
    m = C.c_uint64(0)
    check(lib.gb_edge_list_parse(text, len(text), None, None, None, C.byref(m)))
    src = np.empty(m.value, np.uint32)
    dst = np.empty(m.value, np.uint32)
    w = np.empty(m.value, np.float32) if with_values else None
    check(lib.gb_edge_list_parse(text, len(text), _ptr(src), _ptr(dst), _ptr(w), C.byref(m)))
    return (src, dst, w) if with_values else (src, dst)


def _check_host_csr(off: np.ndarray, tgt: np.ndarray, what: str) -> None:
    """The C side reads off[n] and copies off[n] targets: reject arrays that are too short here."""
    if off.ndim != 1 or len(off) < 2:
        raise ValueError(f"{what} offsets need node_count + 1 >= 2 entries")
    if tgt.ndim != 1 or len(tgt) < int(off[-1]):
        raise ValueError(f"{what} targets hold {len(tgt)} entries but the offsets end at {int(off[-1])}")


def _edges_from_numpy(arr) -> tuple[np.ndarray, np.ndarray]:
    a = np.asarray(arr)
    if a.ndim != 2 or a.shape[1] < 2:
        # crates/mate/src/graphs/mod.rs:441-449
        raise TypeError("Can only create a graph from a 2-dimensional array with at least 2 columns")
    if a.dtype != np.uint32:
        if not np.issubdtype(a.dtype, np.integer) or (a.size and (a.min() < 0 or a.max() > 0xFFFFFFFF)):
            raise TypeError("node ids must be 32-bit unsigned integers")
        a = a.astype(np.uint32)
    return np.ascontiguousarray(a[:, 0]), np.ascontiguousarray(a[:, 1])


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


# ---- graph handles ---------------------------------------------------------------------------
class _Handle:
    """Owns a gb_graph* and a lazily materialised read-only host mirror for neighbor views."""

    def __init__(self, raw_ptr, load_micros=0):
        self._g = raw_ptr
        self.load_micros = int(load_micros)
        info = _capi.GraphInfo()
        check(lib.gb_graph_get_info(self._g, C.byref(info)))
        self._info = info
        self._host = {}

    def __del__(self):
        g, self._g = getattr(self, "_g", None), None
        if g:
            try:
                lib.gb_graph_free(g)
            except Exception:  # interpreter shutdown
                pass

    def _refresh(self):
        check(lib.gb_graph_get_info(self._g, C.byref(self._info)))

    def _mirror(self, which: int):
        """(offsets, targets[, weights]) host copy of one CSR, fetched once (csr.rs:97-117 views)."""
        if which not in self._host:
            n = self._info.node_count
            ln = C.c_uint64(0)
            check(lib.gb_graph_csr_len(self._g, which, C.byref(ln)))
            off = np.empty(n + 1, np.uint32)
            tgt = np.empty(ln.value, np.uint32)
            check(lib.gb_graph_copy_csr(self._g, which, _ptr(off), _ptr(tgt) if ln.value else None, None))
            off.flags.writeable = False
            tgt.flags.writeable = False  # views are read-only (shared_slice.rs:128)
            self._host[which] = (off, tgt)
        return self._host[which]

    def _views_alive(self) -> bool:
        # a numpy view keeps a reference to its base array
        for off, tgt in self._host.values():
            if sys.getrefcount(tgt) > 3:  # tuple entry + loop variable + getrefcount argument
                return True
        return False

    def _row(self, which: int, node: int) -> np.ndarray:
        off, tgt = self._mirror(which)
        n = self._info.node_count
        if not 0 <= node < n:
            raise IndexError(f"node {node} out of range for a graph with {n} nodes")
        return tgt[off[node]:off[node + 1]]

    def _degree(self, which: int, node: int) -> int:
        off, _ = self._mirror(which)
        n = self._info.node_count
        if not 0 <= node < n:
            raise IndexError(f"node {node} out of range for a graph with {n} nodes")
        return int(off[node + 1]) - int(off[node])

    def node_count(self) -> int:
        return int(self._info.node_count)

    def edge_count(self) -> int:
        return int(self._info.edge_count)

    def device_bytes(self) -> int:
        self._refresh()
        return int(self._info.device_bytes)

    def last_timing(self) -> dict:
        t = _capi.Timing()
        check(lib.gb_graph_last_timing(self._g, C.byref(t)))
        return {"total_ms": t.total_ms, "hot_kernel_ms": t.hot_kernel_ms,
                "hot_kernel_launches": int(t.hot_kernel_launches), "kernel_launches": int(t.kernel_launches)}

    def cuda_stream(self) -> int:
        return int(lib.gb_graph_stream(self._g) or 0)

    def __repr__(self):
        return (f"{type(self).__name__} {{ node_count: {self.node_count()}, edge_count: {self.edge_count()}, "
                f"load_took: {_took(self.load_micros)} }}")


def _timed(fn):
    t0 = time.perf_counter()
    out = fn()
    return out, max(1, int((time.perf_counter() - t0) * 1e6))


_PR_MODES = {"auto": _capi.PR_AUTO, "exact": _capi.PR_EXACT, "jacobi": _capi.PR_JACOBI}


class DiGraph(_Handle):
    """A directed graph using 32 bits for node ids — device twin of DirectedCsrGraph<u32>
    (crates/builder/src/graph/csr.rs:364-368; Python surface crates/mate/graph_mate.pyi:46-118)."""

    # -- construction --
    @staticmethod
    def _from_edges(src, dst, weights, node_count, layout) -> "DiGraph":
        out = C.c_void_p()

        def go():
            check(lib.gb_digraph_from_edges_u32(_device, _ptr(src), _ptr(dst), _ptr(weights), len(src),
                                                node_count, _layout_value(layout), C.byref(out)))
        _, micros = _timed(go)
        return DiGraph(out, micros)

    @staticmethod
    def load(path, layout=None, file_format=FileFormat.Graph500) -> "DiGraph":
        """Load a graph from the provided format (crates/mate/src/graphs/digraph.rs:35-44)."""
        t0 = time.perf_counter()
        if file_format is FileFormat.Graph500:
            src, dst, n = _read_graph500(path)
            w = None
        elif file_format is FileFormat.EdgeList:
            src, dst = _read_edge_list(path)
            n, w = 0, None
        else:
            raise TypeError(f"unknown file format {file_format!r}")
        g = DiGraph._from_edges(src, dst, w, n, layout)
        g.load_micros = max(1, int((time.perf_counter() - t0) * 1e6))
        return g

    @staticmethod
    def load_weighted(path, layout=None) -> "DiGraph":
        """Weighted text edge list `<src> <dst> <f32>` (DirectedCsrGraph<u32, (), f32>, for sssp)."""
        src, dst, w = _read_edge_list(path, with_values=True)
        return DiGraph._from_edges(src, dst, w, 0, layout)

    @staticmethod
    def from_numpy(arr, layout=None, weights=None, node_count: int = 0) -> "DiGraph":
        src, dst = _edges_from_numpy(arr)
        w = None if weights is None else np.ascontiguousarray(weights, dtype=np.float32)
        if w is not None and len(w) != len(src):
            raise ValueError("weights must have one entry per edge")
        return DiGraph._from_edges(src, dst, w, node_count, layout)

    @staticmethod
    def from_pandas(df, layout=None) -> "DiGraph":
        return DiGraph.from_numpy(df.to_numpy(), layout)  # crates/mate/src/graphs/mod.rs:169-189

    @staticmethod
    def from_csr(out_offsets, out_targets, in_offsets, in_targets, out_weights=None) -> "DiGraph":
        """Device twin of an already built DirectedCsrGraph (host CSR arrays are uploaded as is)."""
        oo = np.ascontiguousarray(out_offsets, np.uint32)
        ot = np.ascontiguousarray(out_targets, np.uint32)
        io = np.ascontiguousarray(in_offsets, np.uint32)
        it = np.ascontiguousarray(in_targets, np.uint32)
        ow = None if out_weights is None else np.ascontiguousarray(out_weights, np.float32)
        _check_host_csr(oo, ot, "out")
        _check_host_csr(io, it, "in")
        if len(io) != len(oo):
            raise ValueError("in and out offsets must have the same length (node_count + 1)")
        if ow is not None and len(ow) != len(ot):
            raise ValueError("out_weights must have one entry per out target")
        out = C.c_void_p()

        def go():
            check(lib.gb_digraph_from_csr_u32(_device, len(oo) - 1, _ptr(oo), _ptr(ot), _ptr(ow), _ptr(io),
                                              _ptr(it), C.byref(out)))
        _, micros = _timed(go)
        return DiGraph(out, micros)

    @staticmethod
    def for_page_rank(in_offsets, in_targets, out_offsets) -> "DiGraph":
        """Device twin holding only what page_rank reads: the in-CSR and the out-degrees (as out offsets).
        Arrays are used as given (pass pinned uint32 arrays to upload at PCIe speed)."""
        io, it, oo = (np.asarray(a) for a in (in_offsets, in_targets, out_offsets))
        for a in (io, it, oo):
            if a.dtype != np.uint32 or not a.flags.c_contiguous:
                raise TypeError("for_page_rank needs contiguous uint32 arrays")
        _check_host_csr(io, it, "in")
        if len(oo) != len(io):
            raise ValueError("in and out offsets must have the same length (node_count + 1)")
        out = C.c_void_p()

        def go():
            check(lib.gb_digraph_for_page_rank_u32(_device, len(io) - 1, _ptr(io), _ptr(it), _ptr(oo), C.byref(out)))
        _, micros = _timed(go)
        return DiGraph(out, micros)

    @staticmethod
    def rmat(scale: int, edge_factor: int = 16, seed: int = 42, layout=Layout.Sorted, weights=False) -> "DiGraph":
        """Synthetic R-MAT graph generated and built on device (the BASELINE.json workload)."""
        out = C.c_void_p()

        def go():
            check(lib.gb_digraph_rmat(_device, scale, edge_factor, seed, _layout_value(layout), int(bool(weights)),
                                      C.byref(out)))
        _, micros = _timed(go)
        return DiGraph(out, micros)

    # -- accessors --
    def out_degree(self, node: int) -> int:
        return self._degree(_capi.CSR_OUT, node)

    def in_degree(self, node: int) -> int:
        return self._degree(_capi.CSR_IN, node)

    def out_neighbors(self, node: int) -> np.ndarray:
        return self._row(_capi.CSR_OUT, node)

    def in_neighbors(self, node: int) -> np.ndarray:
        return self._row(_capi.CSR_IN, node)

    def copy_out_neighbors(self, node: int) -> list:
        return self._row(_capi.CSR_OUT, node).tolist()

    def copy_in_neighbors(self, node: int) -> list:
        return self._row(_capi.CSR_IN, node).tolist()

    def out_weights(self) -> np.ndarray:
        """f32 edge values aligned with the out-CSR targets (SoA twin of Target<u32, f32>)."""
        ln = C.c_uint64(0)
        check(lib.gb_graph_csr_len(self._g, _capi.CSR_OUT, C.byref(ln)))
        off = np.empty(self.node_count() + 1, np.uint32)
        w = np.empty(ln.value, np.float32)
        check(lib.gb_graph_copy_csr(self._g, _capi.CSR_OUT, _ptr(off), None, _ptr(w)))
        return w

    def csr(self, which: str = "out"):
        """(offsets, targets) host arrays of the out or in CSR (read-only)."""
        return self._mirror(_capi.CSR_OUT if which == "out" else _capi.CSR_IN)

    def to_undirected(self, layout=None) -> "Graph":
        """New, unrelated undirected graph (graph_ops.rs:229; csr.rs:391-464)."""
        out = C.c_void_p()

        def go():
            check(lib.gb_to_undirected(self._g, _layout_value(layout), C.byref(out)))
        _, micros = _timed(go)
        return Graph(out, micros)

    def in_degree_partition(self, parts: int) -> list:
        """graph_ops.rs:431-439 — ranges as [(start, end), ...]"""
        r = np.zeros(parts + 1, np.uint32)
        check(lib.gb_in_degree_partition(self._g, parts, _ptr(r)))
        out = [(int(r[i]), int(r[i + 1])) for i in range(parts) if r[i + 1] > r[i]]
        return out

    def page_rank_plan_info(self) -> dict:
        """Statistics of the device layout the JACOBI page_rank sweeps (built on first use)."""
        st = _capi.PrShardStats()
        check(lib.gb_page_rank_plan_info(self._g, C.byref(st)))
        return st.as_dict()

    # -- algorithms --
    def page_rank(self, *, max_iterations: int = PageRankConfig.DEFAULT_MAX_ITERATIONS,
                  tolerance: float = PageRankConfig.DEFAULT_TOLERANCE,
                  damping_factor: float = PageRankConfig.DEFAULT_DAMPING_FACTOR,
                  mode: str = "auto") -> PageRankResult:
        """page_rank(&graph, PageRankConfig) (page_rank.rs:58-111); keyword-only like
        crates/mate/src/graphs/digraph.rs:126-142.  `mode`: "auto" | "exact" | "jacobi"."""
        cfg = _capi.PageRankConfig(int(max_iterations), float(tolerance), float(damping_factor), _PR_MODES[mode])
        scores = np.empty(self.node_count(), np.float32)
        it, err = C.c_uint64(0), C.c_double(0.0)

        def go():
            check(lib.gb_page_rank(self._g, C.byref(cfg), _ptr(scores), C.byref(it), C.byref(err)))
        _, micros = _timed(go)
        return PageRankResult(scores, int(it.value), float(err.value), micros)

    def wcc(self, *, chunk_size: int = WccConfig.DEFAULT_CHUNK_SIZE,
            neighbor_rounds: int = WccConfig.DEFAULT_NEIGHBOR_ROUNDS,
            sampling_size: int = WccConfig.DEFAULT_SAMPLING_SIZE) -> WccResult:
        """wcc_afforest(&graph, WccConfig).to_vec() (wcc.rs:127-139); keyword-only (digraph.rs:144-160)."""
        cfg = _capi.WccConfig(int(chunk_size), int(neighbor_rounds), int(sampling_size))
        comp = np.empty(self.node_count(), np.uint32)

        def go():
            check(lib.gb_wcc(self._g, C.byref(cfg), _ptr(comp)))
        _, micros = _timed(go)
        return WccResult(comp, micros)

    def wcc_afforest_dss(self, **kw) -> WccResult:
        """wcc_afforest_dss(&graph, config) (wcc.rs:144-156), as `Components::component` reports it: the
        minimum node id of every node's component — the same labels as wcc().  The reference variant
        differs only in its backing union-find (DisjointSetStruct, dss.rs), whose raw `to_vec()` may hold
        non-root ancestors that depend on the thread schedule; nothing consumes it (crates/app/src/app.rs:15
        drops the result), so the device path does not imitate it."""
        return self.wcc(**kw)

    def wcc_baseline(self, **kw) -> WccResult:
        """wcc_baseline(&graph, config) (wcc.rs:103-123): union over every out-edge; same component labels."""
        return self.wcc(**kw)

    def delta_stepping(self, *, start_node: int, delta: float) -> SsspResult:
        """delta_stepping(&graph, DeltaSteppingConfig) (sssp.rs:38-102); needs f32 edge values."""
        if start_node < 0:
            raise ValueError("start_node must be non-negative")
        cfg = _capi.SsspConfig(int(start_node), float(delta))
        dist = np.empty(self.node_count(), np.float32)

        def go():
            check(lib.gb_sssp(self._g, C.byref(cfg), _ptr(dist)))
        _, micros = _timed(go)
        return SsspResult(dist, micros)


class Graph(_Handle):
    """An undirected graph using 32 bits for node ids — device twin of UndirectedCsrGraph<u32>
    (csr.rs:658-661; Python surface graph_mate.pyi:120-168)."""

    @staticmethod
    def _from_edges(src, dst, node_count, layout) -> "Graph":
        out = C.c_void_p()

        def go():
            check(lib.gb_graph_from_edges_u32(_device, _ptr(src), _ptr(dst), len(src), node_count,
                                              _layout_value(layout), C.byref(out)))
        _, micros = _timed(go)
        return Graph(out, micros)

    @staticmethod
    def load(path, layout=None, file_format=FileFormat.Graph500) -> "Graph":
        t0 = time.perf_counter()
        if file_format is FileFormat.Graph500:
            src, dst, n = _read_graph500(path)
        elif file_format is FileFormat.EdgeList:
            src, dst = _read_edge_list(path)
            n = 0
        else:
            raise TypeError(f"unknown file format {file_format!r}")
        g = Graph._from_edges(src, dst, n, layout)
        g.load_micros = max(1, int((time.perf_counter() - t0) * 1e6))
        return g

    @staticmethod
    def from_numpy(arr, layout=None, node_count: int = 0) -> "Graph":
        src, dst = _edges_from_numpy(arr)
        return Graph._from_edges(src, dst, node_count, layout)

    @staticmethod
    def from_pandas(df, layout=None) -> "Graph":
        return Graph.from_numpy(df.to_numpy(), layout)

    @staticmethod
    def from_csr(offsets, targets) -> "Graph":
        off = np.ascontiguousarray(offsets, np.uint32)
        tgt = np.ascontiguousarray(targets, np.uint32)
        _check_host_csr(off, tgt, "undirected")
        out = C.c_void_p()

        def go():
            check(lib.gb_graph_from_csr_u32(_device, len(off) - 1, _ptr(off), _ptr(tgt), C.byref(out)))
        _, micros = _timed(go)
        return Graph(out, micros)

    @staticmethod
    def rmat(scale: int, edge_factor: int = 16, seed: int = 42, layout=Layout.Sorted) -> "Graph":
        out = C.c_void_p()

        def go():
            check(lib.gb_graph_rmat(_device, scale, edge_factor, seed, _layout_value(layout), C.byref(out)))
        _, micros = _timed(go)
        return Graph(out, micros)

    def degree(self, node: int) -> int:
        return self._degree(_capi.CSR_UNDIRECTED, node)

    def neighbors(self, node: int) -> np.ndarray:
        return self._row(_capi.CSR_UNDIRECTED, node)

    def copy_neighbors(self, node: int) -> list:
        return self._row(_capi.CSR_UNDIRECTED, node).tolist()

    def csr(self):
        return self._mirror(_capi.CSR_UNDIRECTED)

    def make_degree_ordered(self) -> None:
        """Relabel by descending degree, in place (graph_ops.rs:173, 511-638)."""
        if self._views_alive():
            # crates/mate/src/graphs/mod.rs:264-276
            raise ValueError("Graph cannot be reordered because there are references to this graph from neighbor lists.")

        def go():
            check(lib.gb_make_degree_ordered(self._g))
        _, micros = _timed(go)
        self._host.clear()
        self.load_micros += micros

    def global_triangle_count(self) -> TriangleCountResult:
        """global_triangle_count(&graph) (triangle_count.rs:22-86)."""
        tri = C.c_uint64(0)

        def go():
            check(lib.gb_triangle_count(self._g, C.byref(tri)))
        _, micros = _timed(go)
        return TriangleCountResult(int(tri.value), micros)
