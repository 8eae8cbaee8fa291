"""The Flight front end (graph_b200/flight.py) against a real in-process Flight client, with the CPU oracle
standing in for the device engine: what is checked here is the reference server's protocol
(crates/server/src/{actions,server,catalog}.rs) — action names, JSON shapes, result documents, record-batch
chunking, error classes — not the kernels (tests/test_zz_flight_gpu.py runs the same calls on the device)."""
import json

import numpy as np
import pytest

import oracle

pa = pytest.importorskip("pyarrow")
fl = pytest.importorskip("pyarrow.flight")


class _G:
    def __init__(self, src, dst, n, weights=None, undirected=False, layout=oracle.UNSORTED):
        self.src, self.dst, self.n, self.w, self.undirected, self.layout = src, dst, n, weights, undirected, layout
        self.relabeled = False

    def node_count(self):
        return self.n

    def edge_count(self):
        return len(self.src)


class OracleEngine:
    """graph_b200.flight's engine interface on top of oracle/ (test infrastructure)."""
    L = {"Unsorted": oracle.UNSORTED, "Sorted": oracle.SORTED, "Deduplicated": oracle.DEDUPLICATED}

    def load(self, path, file_format, orientation, layout):
        raw = open(path, "rb").read()
        if file_format == "Graph500":
            src, dst = oracle.graph500_decode(raw)
            n, w = len(src) // 16, None
        elif file_format == "EdgeListWeighted":
            src, dst, w = oracle.edgelist_parse(raw, with_values=True)
            n = oracle.node_count(src, dst)
        else:
            src, dst = oracle.edgelist_parse(raw)
            n, w = oracle.node_count(src, dst), None
        und = orientation == "Undirected"
        kind = ("undirected" if und else "directed") + ("+weighted" if w is not None else "")
        return kind, _G(src, dst, n, w, und, self.L[layout])

    def from_edges(self, src, dst, orientation, layout):
        und = orientation == "Undirected"
        return ("undirected" if und else "directed"), _G(src, dst, oracle.node_count(src, dst), None, und, self.L[layout])

    def page_rank(self, g, max_iterations, tolerance, damping_factor):
        out = oracle.csr_build(g.src, g.dst, g.n, oracle.OUTGOING, g.layout)
        inc = oracle.csr_build(g.src, g.dst, g.n, oracle.INCOMING, g.layout)
        s, it, err = oracle.page_rank_seq(inc[0], inc[1], out[0], max_iterations, tolerance, damping_factor)
        return s, int(it), float(err)

    def wcc(self, g, chunk_size, neighbor_rounds, sampling_size):
        out = oracle.csr_build(g.src, g.dst, g.n, oracle.OUTGOING, g.layout)
        return oracle.wcc_min_label(out[0], out[1])

    def sssp(self, g, start_node, delta):
        off, tgt, w = oracle.csr_build(g.src, g.dst, g.n, oracle.OUTGOING, g.layout, g.w)
        return oracle.sssp_delta_stepping(off, tgt, w, start_node, delta)

    def triangle_count(self, g):
        off, tgt = oracle.csr_build(g.src, g.dst, g.n, oracle.UNDIRECTED, g.layout)
        if g.relabeled:
            off, tgt, _ = oracle.make_degree_ordered(off, tgt)
        return oracle.triangle_count(off, tgt)

    def make_degree_ordered(self, g):
        g.relabeled = True

    def to_undirected(self, g, layout):
        return _G(g.src, g.dst, g.n, g.w, True, self.L[layout])


@pytest.fixture(scope="module")
def client():
    from graph_b200.flight import GraphFlightServer
    server = GraphFlightServer("grpc://127.0.0.1:0", engine=OracleEngine())
    c = fl.connect(f"grpc://127.0.0.1:{server.port}")
    yield c
    c.close()
    server.shutdown()


def act(client, kind, doc=None):
    body = b"" if doc is None else json.dumps(doc).encode()
    return json.loads(next(client.do_action(fl.Action(kind, body))).body.to_pybytes())


def test_list_actions_names_the_six_reference_actions(client):
    assert [a.type for a in client.list_actions()] == ["create", "list", "remove", "compute", "to_relabeled",
                                                      "to_undirected"]


def test_create_compute_and_stream_the_property(client, golden_dir):
    path = str(golden_dir / "example.el")
    r = act(client, "create", {"graph_name": "ex", "file_format": "EdgeList", "path": path, "csr_layout": "Sorted",
                               "orientation": "Directed"})
    src, dst = oracle.edgelist_parse(open(path, "rb").read())
    assert r["node_count"] == oracle.node_count(src, dst) and r["edge_count"] == len(src) and "create_millis" in r
    infos = act(client, "list")["graph_infos"]
    assert {"graph_name": "ex", "graph_type": "directed", "node_count": r["node_count"],
            "edge_count": r["edge_count"]} in infos
    pr = act(client, "compute", {"graph_name": "ex", "property_key": "pr",
                                 "algorithm": {"PageRank": {"max_iterations": 10, "tolerance": 1e-4,
                                                            "damping_factor": 0.85}}})
    assert pr["property_id"] == {"graph_name": "ex", "property_key": "pr"}
    assert set(pr["algo_result"]) == {"iterations", "error", "compute_millis"}
    table = client.do_get(fl.Ticket(json.dumps(pr["property_id"]).encode())).read_all()
    assert table.schema.names == ["page_rank"] and table.schema.field(0).type == pa.float32()
    out = oracle.csr_build(src, dst, r["node_count"], oracle.OUTGOING, oracle.SORTED)
    inc = oracle.csr_build(src, dst, r["node_count"], oracle.INCOMING, oracle.SORTED)
    want, it, err = oracle.page_rank_seq(inc[0], inc[1], out[0], 10, 1e-4, 0.85)
    assert table.column(0).to_numpy().tobytes() == want.tobytes() and pr["algo_result"]["iterations"] == it
    wcc = act(client, "compute", {"graph_name": "ex", "property_key": "component",
                                  "algorithm": {"Wcc": {"chunk_size": 16384, "neighbor_rounds": 2, "sampling_size": 1024}}})
    comp = client.do_get(fl.Ticket(json.dumps(wcc["property_id"]).encode())).read_all()
    assert comp.schema.names == ["component"] and comp.schema.field(0).type == pa.uint64()
    # algorithms on the wrong kind of graph are invalid arguments, with the reference's messages
    with pytest.raises(pa.ArrowInvalid, match="Triangle count requires an undirected graph"):
        act(client, "compute", {"graph_name": "ex", "property_key": "tc", "algorithm": {"TriangleCount": None}})
    with pytest.raises(pa.ArrowInvalid, match="Sssp requires a directed, weighted graph"):
        act(client, "compute", {"graph_name": "ex", "property_key": "d", "algorithm": {"Sssp": {"start_node": 0, "delta": 3.0}}})
    with pytest.raises(pa.ArrowInvalid, match="Relabelling directed graphs is not supported"):
        act(client, "to_relabeled", {"graph_name": "ex"})
    # to_undirected replaces the catalog entry; triangle count then answers without storing a property
    assert "to_undirected_millis" in act(client, "to_undirected", {"graph_name": "ex", "csr_layout": "Sorted"})
    assert [i["graph_type"] for i in act(client, "list")["graph_infos"] if i["graph_name"] == "ex"] == ["undirected"]
    assert "to_relabeled_millis" in act(client, "to_relabeled", {"graph_name": "ex"})
    tc = act(client, "compute", {"graph_name": "ex", "property_key": "tc", "algorithm": "TriangleCount"})
    off, tgt = oracle.csr_build(src, dst, r["node_count"], oracle.UNDIRECTED, oracle.SORTED)
    assert tc["triangle_count"] == oracle.triangle_count(off, tgt) and "property_id" not in tc
    removed = act(client, "remove", {"graph_name": "ex"})
    assert removed["graph_name"] == "ex" and removed["graph_type"] == "undirected"


def test_put_streams_edges_and_properties_come_back_in_10k_batches(client):
    src, dst = oracle.rmat_edges(15, seed=3)
    n = int(max(src.max(), dst.max())) + 1
    table = pa.table({"source": pa.array(src.astype(np.int64)), "target": pa.array(dst.astype(np.int64))})
    cmd = {"graph_name": "put", "edge_count": len(src), "csr_layout": "Sorted", "orientation": "Directed"}
    writer, reader = client.do_put(fl.FlightDescriptor.for_command(json.dumps(cmd).encode()), table.schema)
    writer.write_table(table, max_chunksize=100_000)
    writer.done_writing()
    result = json.loads(reader.read().to_pybytes())
    writer.close()
    assert result["node_count"] == n and result["edge_count"] == len(src)
    pr = act(client, "compute", {"graph_name": "put", "property_key": "pr",
                                 "algorithm": {"PageRank": {"max_iterations": 5, "tolerance": 0.0, "damping_factor": 0.85}}})
    rd = client.do_get(fl.Ticket(json.dumps(pr["property_id"]).encode()))
    sizes = [c.data.num_rows for c in rd]
    assert sum(sizes) == n and max(sizes) == 10_000 and all(s == 10_000 for s in sizes[:-1])


def test_weighted_edge_list_serves_sssp(client, tmp_path):
    p = tmp_path / "w.el"
    p.write_text("0 1 0.5\n1 2 1.25\n0 2 4.0\n2 3 1.0\n")
    r = act(client, "create", {"graph_name": "w", "file_format": "EdgeListWeighted", "path": str(p)})
    assert r["node_count"] == 4 and r["edge_count"] == 4
    assert [i["graph_type"] for i in act(client, "list")["graph_infos"] if i["graph_name"] == "w"] == ["directed+weighted"]
    s = act(client, "compute", {"graph_name": "w", "property_key": "dist", "algorithm": {"Sssp": {"start_node": 0, "delta": 1.0}}})
    t = client.do_get(fl.Ticket(json.dumps(s["property_id"]).encode())).read_all()
    assert t.schema.names == ["distance"] and t.column(0).to_pylist() == [0.0, 0.5, 1.75, 2.75]


def test_errors_keep_their_classes(client):
    with pytest.raises(pa.ArrowInvalid, match="Unknown action type: nope"):
        act(client, "nope")
    with pytest.raises((KeyError, pa.ArrowKeyError), match="Graph with name 'ghost' not found"):
        act(client, "remove", {"graph_name": "ghost"})
    with pytest.raises((KeyError, pa.ArrowKeyError), match="Property Id not found"):
        client.do_get(fl.Ticket(json.dumps({"graph_name": "ghost", "property_key": "x"}).encode())).read_all()
    with pytest.raises(fl.FlightInternalError, match="JsonError"):
        act(client, "create", {"graph_name": "x"})                      # missing fields
    with pytest.raises(fl.FlightInternalError, match="JsonError"):
        client.do_action(fl.Action("compute", b"not json")).__next__()
    with pytest.raises(fl.FlightInternalError, match="unknown variant"):
        act(client, "create", {"graph_name": "x", "file_format": "Parquet", "path": "/nope"})
    with pytest.raises(pa.ArrowInvalid, match="node ids must fit u32"):
        t = pa.table({"s": pa.array([1 << 33], pa.int64()), "t": pa.array([0], pa.int64())})
        cmd = {"graph_name": "big", "edge_count": 1}
        w, r = client.do_put(fl.FlightDescriptor.for_command(json.dumps(cmd).encode()), t.schema)
        w.write_table(t)
        w.done_writing()
        r.read()
        w.close()


def test_device_engine_calls_bind_to_the_package_api():
    """B200Engine is exercised on the GPU only; here its calls are bound against the real signatures of
    graph_b200 (the library loads without a device), so that a renamed keyword cannot hide until then."""
    import inspect
    import graph_b200 as gb
    from graph_b200.flight import B200Engine, LAYOUTS
    eng = B200Engine()
    for name in LAYOUTS:
        assert eng._layout(name) is getattr(gb.Layout, name)
    lay, fmt = gb.Layout.Sorted, gb.FileFormat.EdgeList
    assert gb.FileFormat.Graph500 is not gb.FileFormat.EdgeList
    bind = lambda fn, *a, **k: inspect.signature(fn).bind(*a, **k)
    bind(gb.DiGraph.load, "p", lay, fmt)
    bind(gb.DiGraph.load_weighted, "p", lay)
    bind(gb.Graph.load, "p", lay, fmt)
    bind(gb.DiGraph.from_numpy, np.zeros((1, 2), np.uint32), lay)
    bind(gb.Graph.from_numpy, np.zeros((1, 2), np.uint32), lay)
    bind(gb._read_edge_list, "p", with_values=True)
    bind(gb.DiGraph.page_rank, None, max_iterations=1, tolerance=0.0, damping_factor=0.85)
    bind(gb.DiGraph.wcc, None, chunk_size=1, neighbor_rounds=1, sampling_size=1)
    bind(gb.DiGraph.delta_stepping, None, start_node=0, delta=1.0)
    bind(gb.DiGraph.to_undirected, None, lay)
    bind(gb.Graph.global_triangle_count, None)
    bind(gb.Graph.make_degree_ordered, None)
    pr = gb.PageRankResult(np.zeros(2, np.float32), 3, 0.5, 1)
    assert pr.scores().dtype == np.float32 and pr.ran_iterations == 3 and pr.error == 0.5
    assert gb.WccResult(np.zeros(2, np.uint32), 1).components().dtype == np.uint32
    assert gb.SsspResult(np.zeros(2, np.float32), 1).distances().dtype == np.float32
    assert gb.TriangleCountResult(7, 1).triangles == 7
    # read-only result arrays (the package marks them so) convert to Arrow without a copy error
    assert pa.array(pr.scores(), pa.float32()).to_pylist() == [0.0, 0.0]
