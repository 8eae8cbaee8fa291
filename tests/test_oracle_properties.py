"""Property tests of the CPU oracle (hypothesis): every restated algorithm is cross-checked against an
independent statement of its result on random multigraphs with self loops, duplicates and isolated
vertices.  Parity of the CUDA path hinges on the oracle, so the oracle is not trusted on goldens alone."""
import numpy as np
from hypothesis import given, settings, strategies as st

import oracle


@st.composite
def edge_lists(draw, max_n=40, max_m=160):
    n = draw(st.integers(1, max_n))
    m = draw(st.integers(0, max_m))
    src = np.array(draw(st.lists(st.integers(0, n - 1), min_size=m, max_size=m)), dtype=np.uint32)
    dst = np.array(draw(st.lists(st.integers(0, n - 1), min_size=m, max_size=m)), dtype=np.uint32)
    return n, src, dst


SETTINGS = dict(max_examples=60, deadline=None)


@settings(**SETTINGS)
@given(edge_lists())
def test_csr_layouts_are_consistent(g):
    n, src, dst = g
    for direction in (oracle.OUTGOING, oracle.INCOMING, oracle.UNDIRECTED):
        uo, ut = oracle.csr_build(src, dst, n, direction, oracle.UNSORTED)
        so, stg = oracle.csr_build(src, dst, n, direction, oracle.SORTED)
        do, dt = oracle.csr_build(src, dst, n, direction, oracle.DEDUPLICATED)
        assert (uo == so).all() and uo[0] == 0 and uo[-1] == len(ut)
        assert len(ut) == (2 * len(src) if direction == oracle.UNDIRECTED else len(src))
        for v in range(n):
            row_u = ut[uo[v]:uo[v + 1]]
            row_s = stg[so[v]:so[v + 1]]
            row_d = dt[do[v]:do[v + 1]]
            assert sorted(row_u.tolist()) == row_s.tolist()                       # csr.rs:886-895
            assert row_d.tolist() == sorted(set(row_s.tolist()) - {v})            # csr.rs:897-948
    # unsorted rows keep the edge-list order (single-thread build)
    uo, ut = oracle.csr_build(src, dst, n, oracle.OUTGOING, oracle.UNSORTED)
    for v in range(n):
        assert ut[uo[v]:uo[v + 1]].tolist() == dst[src == v].tolist()


@settings(**SETTINGS)
@given(edge_lists())
def test_make_degree_ordered_is_the_documented_permutation(g):
    n, src, dst = g
    off, tgt = oracle.csr_build(src, dst, n, oracle.UNDIRECTED, oracle.SORTED)
    noff, ntgt, nid = oracle.make_degree_ordered(off, tgt)
    deg = np.diff(off.astype(np.int64))
    order = sorted(range(n), key=lambda v: (-deg[v], -v))            # (degree, id) descending, graph_ops.rs:555
    assert [int(nid[v]) for v in order] == list(range(n))
    assert np.diff(noff.astype(np.int64)).tolist() == [int(deg[v]) for v in order]
    for v in range(n):
        want = sorted(int(nid[t]) for t in tgt[off[v]:off[v + 1]])
        assert ntgt[noff[nid[v]]:noff[nid[v] + 1]].tolist() == want


@settings(**SETTINGS)
@given(edge_lists(), st.integers(0, 4), st.integers(0, 64), st.integers(0, 2 ** 32), st.sampled_from([1, 3]))
def test_afforest_equals_min_label(g, rounds, samples, seed, threads):
    n, src, dst = g
    oo, ot = oracle.csr_build(src, dst, n, oracle.OUTGOING, oracle.SORTED)
    io, it = oracle.csr_build(src, dst, n, oracle.INCOMING, oracle.SORTED)
    want = oracle.wcc_min_label(oo, ot)
    got = oracle.wcc_afforest(oo, ot, io, it, neighbor_rounds=rounds, sampling_size=samples, rng_seed=seed,
                              threads=threads)
    assert (got == want).all()
    # independent statement: union-find in numpy-free python
    parent = list(range(n))
    def find(x):
        while parent[x] != x:
            parent[x] = parent[parent[x]]
            x = parent[x]
        return x
    for a, b in zip(src.tolist(), dst.tolist()):
        ra, rb = find(a), find(b)
        if ra != rb:
            parent[max(ra, rb)] = min(ra, rb)
    roots = [find(v) for v in range(n)]
    mins = {}
    for v, r in enumerate(roots):
        mins[r] = min(mins.get(r, v), v)
    assert want.tolist() == [mins[r] for r in roots]


@settings(**SETTINGS)
@given(edge_lists(), st.floats(0.01, 50.0), st.integers(0, 2 ** 31))
def test_delta_stepping_equals_label_correcting_fixed_point(g, delta, seed):
    n, src, dst = g
    rng = np.random.default_rng(seed)
    w = rng.integers(0, 1 << 12, len(src)).astype(np.float32) / np.float32(64.0)   # includes zero weights
    off, tgt, ww = oracle.csr_build(src, dst, n, oracle.OUTGOING, oracle.SORTED, w)
    start = int(rng.integers(0, n))
    a = oracle.sssp_delta_stepping(off, tgt, ww, start, float(np.float32(delta)))
    b = oracle.sssp_bellman_ford(off, tgt, ww, start)
    assert a.tobytes() == b.tobytes()
    # Dijkstra in f32 (heap) reaches the same fixed point
    import heapq
    dist = [np.float32(np.finfo(np.float32).max)] * n
    dist[start] = np.float32(0)
    heap = [(0.0, start)]
    while heap:
        d, u = heapq.heappop(heap)
        if np.float32(d) > dist[u]:
            continue
        for e in range(off[u], off[u + 1]):
            nd = np.float32(dist[u] + ww[e])
            if nd < dist[tgt[e]]:
                dist[tgt[e]] = nd
                heapq.heappush(heap, (float(nd), int(tgt[e])))
    assert [float(x) for x in dist] == a.tolist()


@settings(**SETTINGS)
@given(edge_lists(max_n=24, max_m=120))
def test_triangle_count_matches_the_documented_sum(g):
    n, src, dst = g
    off, tgt = oracle.csr_build(src, dst, n, oracle.UNDIRECTED, oracle.SORTED)
    rows = [tgt[off[v]:off[v + 1]].tolist() for v in range(n)]
    want = 0
    for u in range(n):
        su = set(rows[u])
        for v in rows[u]:
            if v > u:
                break
            for w in rows[v]:
                if w > v:
                    break
                want += w in su
    assert oracle.triangle_count(off, tgt) == want
    assert oracle.triangle_count(off, tgt, threads=3) == want
    # on a Deduplicated graph the sum is the number of triangles and is relabelling invariant
    doff, dtgt = oracle.csr_build(src, dst, n, oracle.UNDIRECTED, oracle.DEDUPLICATED)
    adj = [set(dtgt[doff[v]:doff[v + 1]].tolist()) for v in range(n)]
    tri = sum(1 for a in range(n) for b in adj[a] if b < a for c in adj[b] if c < b and c in adj[a])
    assert oracle.triangle_count(doff, dtgt) == tri
    noff, ntgt, _ = oracle.make_degree_ordered(doff, dtgt)
    assert oracle.triangle_count(noff, ntgt) == tri


@settings(**SETTINGS)
@given(edge_lists(), st.integers(1, 12), st.floats(0.0, 1.0))
def test_page_rank_schedules(g, iters, damping):
    n, src, dst = g
    oo, _ = oracle.csr_build(src, dst, n, oracle.OUTGOING, oracle.SORTED)
    io, it = oracle.csr_build(src, dst, n, oracle.INCOMING, oracle.SORTED)
    d = float(np.float32(damping))
    seq, k, err = oracle.page_rank_seq(io, it, oo, iters, 0.0, d)
    mt, k2, err2 = oracle.page_rank_mt(io, it, oo, iters, 0.0, d, threads=3)   # n <= 16384: one chunk, one thread
    assert k == k2 == iters and seq.tobytes() == mt.tobytes() and err == err2
    # python restatement of page_rank.rs:58-168 in f32, in place
    f = np.float32
    nf = f(n)
    init, base = f(1) / nf, (f(1) - f(d)) / nf
    outdeg = np.diff(oo.astype(np.int64)).astype(np.float32)
    with np.errstate(divide="ignore"):
        out = (init / outdeg).astype(np.float32)
    sc = np.full(n, init, dtype=np.float32)
    for _ in range(iters):
        for u in range(n):
            tot = f(0)
            for e in range(io[u], io[u + 1]):
                tot = f(tot + out[it[e]])
            new = f(base + f(f(d) * tot))
            sc[u] = new
            with np.errstate(divide="ignore"):
                out[u] = f(new / outdeg[u])
    assert sc.tobytes() == seq.tobytes()
    # Jacobi with f32 sums == Jacobi with f64 sums up to summation rounding
    j32, _, _ = oracle.page_rank_jacobi(io, it, oo, iters, 0.0, d, acc64=False)
    j64, _, _ = oracle.page_rank_jacobi(io, it, oo, iters, 0.0, d, acc64=True)
    assert np.allclose(j32, j64, rtol=1e-5, atol=0)


@settings(**SETTINGS)
@given(edge_lists(), st.integers(1, 6))
def test_in_degree_partition_covers_and_balances(g, parts):
    n, src, dst = g
    io, _ = oracle.csr_build(src, dst, n, oracle.INCOMING, oracle.SORTED)
    r = oracle.in_degree_partition(io, parts).tolist()
    assert r[0] == 0 and r[-1] == n and r == sorted(r) and len(r) - 1 <= parts
    batch = -(-len(src) // parts)
    for a, b in zip(r[:-2], r[1:-1]):                      # every closed range reached the batch size
        assert int(io[b]) - int(io[a]) >= batch
