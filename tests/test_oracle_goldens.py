"""Pins the CPU oracle (oracle/oracle.c) against every golden the reference's own tests hold for
the hot path (tests/golden/reference_goldens.json, transcribed with file:line citations).
CPU-only: runs in the build container."""
import numpy as np
import pytest

import oracle
from conftest import LAYOUTS, edges_to_arrays


def f32(strings):
    return np.array([np.float32(s) for s in strings], dtype=np.float32)


def digraph(edges, layout, n=None):
    src, dst, w = edges_to_arrays(edges)
    n = n or oracle.node_count(src, dst)
    out = oracle.csr_build(src, dst, n, oracle.OUTGOING, layout, w)
    inc = oracle.csr_build(src, dst, n, oracle.INCOMING, layout, w)
    return n, out, inc


def test_page_rank_13_nodes_bit_exact(goldens):
    g = goldens["page_rank_13_nodes"]
    n, out, inc = digraph(g["edges"], LAYOUTS[g["layout"]])
    cfg = g["config"]
    scores, it, err = oracle.page_rank_seq(inc[0], inc[1], out[0], cfg["max_iterations"],
                                           cfg["tolerance"], cfg["damping_factor"])
    assert it == g["iterations"]
    assert scores.tobytes() == f32(g["scores"]).tobytes()
    # the multi-threaded restatement degenerates to one chunk here (n <= 16384): same bits
    s2, it2, _ = oracle.page_rank_mt(inc[0], inc[1], out[0], cfg["max_iterations"],
                                     cfg["tolerance"], cfg["damping_factor"], threads=4)
    assert it2 == it and s2.tobytes() == scores.tobytes()


def test_page_rank_two_components_bit_exact(goldens):
    g = goldens["page_rank_two_components"]
    n, out, inc = digraph(g["edges"], LAYOUTS[g["layout"]])
    scores, it, err = oracle.page_rank_seq(inc[0], inc[1], out[0])
    assert scores.tobytes() == f32(g["scores"]).tobytes()


def test_page_rank_jacobi_differs_from_reference_schedule(goldens):
    """Documents SURVEY.md's finding: the reference is an in-place sweep, NOT Jacobi."""
    g = goldens["page_rank_13_nodes"]
    n, out, inc = digraph(g["edges"], LAYOUTS[g["layout"]])
    seq, _, _ = oracle.page_rank_seq(inc[0], inc[1], out[0], 10, 1e-4, 0.85)
    jac, _, _ = oracle.page_rank_jacobi(inc[0], inc[1], out[0], 10, 1e-4, 0.85)
    assert np.max(np.abs(seq - jac) / seq) > 1e-3
    # ... but both schedules share the fixed point
    seq, _, _ = oracle.page_rank_seq(inc[0], inc[1], out[0], 500, 0.0, 0.85)
    jac, _, _ = oracle.page_rank_jacobi(inc[0], inc[1], out[0], 500, 0.0, 0.85)
    assert np.max(np.abs(seq - jac) / seq) < 2e-6


def test_page_rank_scale8_properties(goldens, scale8_edges):
    src, dst, n = scale8_edges
    out = oracle.csr_build(src, dst, n, oracle.OUTGOING, oracle.SORTED)
    inc = oracle.csr_build(src, dst, n, oracle.INCOMING, oracle.SORTED)
    s, it, err = oracle.page_rank_seq(inc[0], inc[1], out[0])
    assert it >= 1 and err < 1.0 and (s > 0).all() and len(s) == 256
    assert oracle.page_rank_seq(inc[0], inc[1], out[0], max_iterations=1)[1] == 1
    assert oracle.page_rank_seq(inc[0], inc[1], out[0], tolerance=1.0)[1] == 1
    s0, it0, _ = oracle.page_rank_seq(inc[0], inc[1], out[0], damping=0.0)
    assert it0 == 1 and (s0 == np.float32(goldens["page_rank_scale8_properties"]["damping_zero_score"])).all()


def test_sssp_golden(goldens):
    g = goldens["sssp"]
    src, dst, w = edges_to_arrays(g["edges"])
    n = oracle.node_count(src, dst)
    off, tgt, ww = oracle.csr_build(src, dst, n, oracle.OUTGOING, LAYOUTS[g["layout"]], w)
    d = oracle.sssp_delta_stepping(off, tgt, ww, g["start_node"], g["delta"])
    assert d.tolist() == g["distances"]
    assert oracle.sssp_bellman_ford(off, tgt, ww, g["start_node"]).tolist() == g["distances"]


def test_triangle_count_goldens(goldens, scale8_edges):
    for g in goldens["triangle_count"]:
        src, dst, _ = edges_to_arrays(g["edges"])
        n = oracle.node_count(src, dst)
        off, tgt = oracle.csr_build(src, dst, n, oracle.UNDIRECTED, LAYOUTS[g["layout"]])
        assert oracle.triangle_count(off, tgt) == g["triangles"], g["cite"]
    src, dst, n = scale8_edges
    off, tgt = oracle.csr_build(src, dst, n, oracle.UNDIRECTED, oracle.SORTED)
    sd = goldens["survey_derived"]
    assert oracle.triangle_count(off, tgt) == sd["scale8_triangles_sorted_unrelabelled"]
    noff, ntgt, _ = oracle.make_degree_ordered(off, tgt)
    want = goldens["triangle_count_scale8_degree_ordered"]["triangles"]
    assert oracle.triangle_count(noff, ntgt) == want
    assert oracle.triangle_count(noff, ntgt, threads=4) == want
    doff, dtgt = oracle.csr_build(src, dst, n, oracle.UNDIRECTED, oracle.DEDUPLICATED)
    assert oracle.triangle_count(doff, dtgt) == sd["scale8_triangles_deduplicated"]


def test_scale8_lists(goldens, scale8_edges):
    g = goldens["scale8_lists"]
    src, dst, n = scale8_edges
    assert n == g["node_count"] and len(src) == g["edge_count"]
    assert int((src == dst).sum()) == goldens["survey_derived"]["scale8_self_loops"]
    ooff, otgt = oracle.csr_build(src, dst, n, oracle.OUTGOING, oracle.SORTED)
    ioff, itgt = oracle.csr_build(src, dst, n, oracle.INCOMING, oracle.SORTED)
    uoff, utgt = oracle.csr_build(src, dst, n, oracle.UNDIRECTED, oracle.SORTED)
    assert otgt[ooff[0]:ooff[1]].tolist() == g["out_neighbors_0"]
    assert itgt[ioff[0]:ioff[1]].tolist() == g["in_neighbors_0"]
    assert utgt[uoff[0]:uoff[1]].tolist() == g["neighbors_0"]
    assert len(utgt) // 2 == g["edge_count"]  # UndirectedCsrGraph::edge_count, csr.rs:687-689


def test_sort_and_deduplicate(goldens):
    g = goldens["sort_and_deduplicate"]
    # rebuild the same rows from an edge list: row r holds targets[off[r]:off[r+1]]
    off, tg = g["offsets"], g["targets"]
    edges = [(r, t) for r in range(len(off) - 1) for t in tg[off[r]:off[r + 1]]]
    src, dst, _ = edges_to_arrays(edges)
    noff, ntgt = oracle.csr_build(src, dst, len(off) - 1 + 4, oracle.OUTGOING, oracle.DEDUPLICATED)
    assert noff[:len(off)].tolist() == g["new_offsets"]
    assert ntgt.tolist() == g["new_targets"]


def test_to_undirected_single_thread_order(goldens):
    g = goldens["to_undirected"]
    src, dst, _ = edges_to_arrays(g["edges"])
    n = oracle.node_count(src, dst)
    # to_undirected feeds the out-CSR rows as an edge list (csr.rs:391-464); for this input the
    # out-CSR order of an Unsorted single-thread build is the edge-list order grouped by source.
    ooff, otgt = oracle.csr_build(src, dst, n, oracle.OUTGOING, oracle.UNSORTED)
    s2 = np.repeat(np.arange(n, dtype=np.uint32), np.diff(ooff.astype(np.int64)))
    for name, want in g["neighbors_0"].items():
        off, tgt = oracle.csr_build(s2, otgt, n, oracle.UNDIRECTED, LAYOUTS[name])
        assert tgt[off[0]:off[1]].tolist() == want, name


def test_to_undirected_layouts(goldens):
    g = goldens["to_undirected_layouts"]
    src, dst, _ = edges_to_arrays(g["edges"])
    n = oracle.node_count(src, dst)
    for name in ("Sorted", "Deduplicated"):
        off, tgt = oracle.csr_build(src, dst, n, oracle.UNDIRECTED, LAYOUTS[name])
        got = [tgt[off[v]:off[v + 1]].tolist() for v in range(n)]
        assert got == g[name]


def test_relabel_by_degree(goldens):
    g = goldens["relabel_by_degree"]
    src, dst, _ = edges_to_arrays(g["edges"])
    n = oracle.node_count(src, dst)
    off, tgt = oracle.csr_build(src, dst, n, oracle.UNDIRECTED, oracle.UNSORTED)
    noff, ntgt, nid = oracle.make_degree_ordered(off, tgt)
    assert nid.tolist() == g["new_id"]
    assert np.diff(noff.astype(np.int64)).tolist() == g["degrees"]
    assert [ntgt[noff[v]:noff[v + 1]].tolist() for v in range(n)] == g["neighbors"]


def test_greedy_partition(goldens):
    g = goldens["greedy_partition"]
    w = np.array(g["weights"], dtype=np.uint32)
    off = np.concatenate([[0], np.cumsum(w)]).astype(np.uint32)
    # in_degree_partition derives batch = ceil(m / parts); pick m so that batch == 6 with 3 parts
    import ctypes as C
    ranges = np.zeros(g["max_batches"] + 1, np.uint32)
    cnt = oracle.lib().orc_in_degree_partition(off, len(w), 6 * 3, 3, ranges)
    assert cnt == 3 and ranges.tolist() == g["ranges"]


def test_afforest_union_and_wcc(goldens, scale8_edges):
    g = goldens["afforest_union"]
    # unions as a graph: edges (u, v) -> result = min label
    src, dst, _ = edges_to_arrays(g["unions"])
    n = g["size"]
    ooff, otgt = oracle.csr_build(src, dst, n, oracle.OUTGOING, oracle.UNSORTED)
    ioff, itgt = oracle.csr_build(src, dst, n, oracle.INCOMING, oracle.UNSORTED)
    comp = oracle.wcc_afforest(ooff, otgt, ioff, itgt)
    assert comp[9] == g["find_9"]
    g = goldens["wcc_two_components"]
    src, dst, _ = edges_to_arrays(g["edges"])
    ooff, otgt = oracle.csr_build(src, dst, 4, oracle.OUTGOING, oracle.UNSORTED)
    ioff, itgt = oracle.csr_build(src, dst, 4, oracle.INCOMING, oracle.UNSORTED)
    comp = oracle.wcc_afforest(ooff, otgt, ioff, itgt)
    for a, b in g["same"]:
        assert comp[a] == comp[b]
    for a, b in g["different"]:
        assert comp[a] != comp[b]
    # scale_8: label == min id of the component, for every config / thread count / sample seed
    src, dst, n = scale8_edges
    ooff, otgt = oracle.csr_build(src, dst, n, oracle.OUTGOING, oracle.SORTED)
    ioff, itgt = oracle.csr_build(src, dst, n, oracle.INCOMING, oracle.SORTED)
    want = oracle.wcc_min_label(ooff, otgt)
    sd = goldens["survey_derived"]
    labels, counts = np.unique(want, return_counts=True)
    assert len(labels) == sd["scale8_components"] and counts.max() == sd["scale8_largest_component"]
    for threads in (1, 4):
        for rounds in (0, 1, 2, 5):
            for seed in (1, 42):
                got = oracle.wcc_afforest(ooff, otgt, ioff, itgt, neighbor_rounds=rounds,
                                          rng_seed=seed, threads=threads)
                assert (got == want).all()


def test_edge_list_files(goldens, golden_dir):
    g = goldens["edge_list_test_el"]
    src, dst = oracle.edgelist_parse((golden_dir / g["file"]).read_bytes())
    n = oracle.node_count(src, dst)
    assert n == g["node_count"] and len(src) == g["edge_count"]
    off, tgt = oracle.csr_build(src, dst, n, oracle.OUTGOING, oracle.SORTED)
    assert [tgt[off[v]:off[v + 1]].tolist() for v in range(n)] == g["out_neighbors"]
    off, tgt = oracle.csr_build(src, dst, n, oracle.UNDIRECTED, oracle.SORTED)
    assert [tgt[off[v]:off[v + 1]].tolist() for v in range(n)] == g["neighbors"]
    # CRLF input (resources/windows.el) and weighted input (resources/test.wel)
    s2, d2 = oracle.edgelist_parse((golden_dir / "windows.el").read_bytes())
    assert len(s2) == 3 or len(s2) > 0
    s3, d3, w3 = oracle.edgelist_parse((golden_dir / "test.wel").read_bytes(), with_values=True)
    assert (s3 == src).all() and (d3 == dst).all()
    assert w3.tolist() == [np.float32(x) for x in (0.1, 0.2, 0.3, 0.4, 0.5, 0.6)]


def test_rmat_stream_is_deterministic_and_in_range():
    s, d = oracle.rmat_edges(10, seed=42)
    assert len(s) == 16 << 10 and s.max() < 1024 and d.max() < 1024
    s2, d2 = oracle.rmat_edges(10, seed=42, first=100, count=50)
    assert (s2 == s[100:150]).all() and (d2 == d[100:150]).all()
    # skew sanity: R-MAT a=.57 puts far more mass on a few vertices than a uniform stream would
    deg = np.bincount(d, minlength=1024)
    assert deg.max() > 20 * deg.mean()
    # the scramble is a bijection of the id space
    import ctypes as C
    ids = np.arange(1024, dtype=np.uint32)
    assert len(np.unique(np.concatenate([s, d]))) > 300


# ---- DisjointSetStruct (crates/algos/src/dss.rs) -----------------------------------------------------
def test_dss_reference_unit_tests():
    """dss.rs:183-220: test_union and test_union_with_path_halving, literally."""
    parents, finds = oracle.dss_ops(10, [(9, 7)])
    assert finds[9] == 7 and finds[7] == 7                                  # dss.rs:187-188
    for pairs, want in (([(9, 7), (7, 4)], 4), ([(9, 7), (7, 4), (4, 2)], 2), ([(9, 7), (7, 4), (4, 2), (2, 0)], 0)):
        assert oracle.dss_ops(10, pairs)[1][9] == want                       # dss.rs:189-194
    chain = [(4, 3), (3, 2), (2, 1), (1, 0), (9, 8), (8, 7), (7, 6), (6, 5)]
    _, finds = oracle.dss_ops(10, chain)
    assert finds[4] == 0 and finds[9] == 5                                   # dss.rs:210-211
    _, finds = oracle.dss_ops(10, chain + [(5, 4)])
    assert (finds == 0).all()                                                # dss.rs:213-217
    # doc tests: union(2, 4) -> find(2) == find(4) == 2 (dss.rs:34-36); union(4, 2) -> find(4) == 2 (:73-75)
    assert oracle.dss_ops(10, [(2, 4)])[1][[2, 4]].tolist() == [2, 2]
    assert oracle.dss_ops(10, [(4, 2)])[1][4] == 2
    # test_union_parallel's postconditions (dss.rs:222-263) on one thread
    pairs = [(i, i + 1) for i in range(500)] + [(i, i + 1) for i in range(501, 999)]
    _, finds = oracle.dss_ops(1000, pairs)
    assert (finds[:501] == finds[0]).all() and finds[500] != finds[501] and (finds[501:] == finds[501]).all()


@pytest.mark.parametrize("variant", ["baseline", "afforest_dss"])
def test_wcc_dss_variants_agree_with_afforest(variant):
    """wcc_baseline / wcc_afforest_dss (wcc.rs:103-156): component(i) is the minimum node id of i's
    component — the same labels as wcc_afforest(..).to_vec(); the DSS to_vec() itself may hold non-root
    ancestors (path halving does not fully compress), which is why it is not a result anybody compares."""
    src, dst = oracle.rmat_edges(12, seed=5)
    n = 1 << 12
    oo, ot = oracle.csr_build(src, dst, n, oracle.OUTGOING, oracle.SORTED)
    io, it = oracle.csr_build(src, dst, n, oracle.INCOMING, oracle.SORTED)
    to_vec, comp = oracle.wcc_dss(oo, ot, io, it, variant)
    want = oracle.wcc_min_label(oo, ot)
    assert (comp == want).all()
    assert (to_vec <= np.arange(n)).all()            # union by min: parents never point upwards
    assert (comp[to_vec] == comp).all()              # every stored parent lies in the node's own component
