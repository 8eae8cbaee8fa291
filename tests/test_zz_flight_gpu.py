"""The Flight front end with its default engine (the device twin): the calls of the reference's example
clients (crates/server/examples/*.py), results checked against the oracle.  Runs last in the GPU suite
(file name): it starts a gRPC server inside the test process."""
import json

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu

pa = pytest.importorskip("pyarrow")
fl = pytest.importorskip("pyarrow.flight")


@pytest.fixture(scope="module")
def client():
    from graph_b200.flight import GraphFlightServer
    server = GraphFlightServer("grpc://127.0.0.1:0")
    c = fl.connect(f"grpc://127.0.0.1:{server.port}")
    yield c
    c.close()
    server.shutdown()


def act(client, kind, doc=None):
    body = b"" if doc is None else json.dumps(doc).encode()
    return json.loads(next(client.do_action(fl.Action(kind, body))).body.to_pybytes())


def prop(client, property_id):
    return client.do_get(fl.Ticket(json.dumps(property_id).encode())).read_all()


def test_flight_compute_on_the_device_matches_the_oracle(client, golden_dir, tmp_path):
    path = str(golden_dir / "example.el")
    src, dst = oracle.edgelist_parse(open(path, "rb").read())
    n = oracle.node_count(src, dst)
    r = act(client, "create", {"graph_name": "ex", "file_format": "EdgeList", "path": path, "csr_layout": "Sorted",
                               "orientation": "Directed"})
    assert (r["node_count"], r["edge_count"]) == (n, len(src))
    out = oracle.csr_build(src, dst, n, oracle.OUTGOING, oracle.SORTED)
    inc = oracle.csr_build(src, dst, n, oracle.INCOMING, oracle.SORTED)
    pr = act(client, "compute", {"graph_name": "ex", "property_key": "pr",
                                 "algorithm": {"PageRank": {"max_iterations": 10, "tolerance": 1e-4,
                                                            "damping_factor": 0.85}}})
    want, it, err = oracle.page_rank_seq(inc[0], inc[1], out[0], 10, 1e-4, 0.85)
    assert pr["algo_result"]["iterations"] == it and pr["algo_result"]["error"] == err
    assert prop(client, pr["property_id"]).column(0).to_numpy().tobytes() == want.tobytes()
    wcc = act(client, "compute", {"graph_name": "ex", "property_key": "component",
                                  "algorithm": {"Wcc": {"chunk_size": 16384, "neighbor_rounds": 2, "sampling_size": 1024}}})
    comp = prop(client, wcc["property_id"]).column(0).to_numpy()
    assert comp.dtype == np.uint64 and (comp == oracle.wcc_min_label(out[0], out[1])).all()
    with pytest.raises(pa.ArrowInvalid, match="Triangle count requires an undirected graph"):
        act(client, "compute", {"graph_name": "ex", "property_key": "tc", "algorithm": {"TriangleCount": None}})
    act(client, "to_undirected", {"graph_name": "ex", "csr_layout": "Sorted"})
    act(client, "to_relabeled", {"graph_name": "ex"})
    tc = act(client, "compute", {"graph_name": "ex", "property_key": "tc", "algorithm": {"TriangleCount": None}})
    off, tgt = oracle.csr_build(src, dst, n, oracle.UNDIRECTED, oracle.SORTED)
    assert tc["triangle_count"] == oracle.triangle_count(off, tgt)
    assert act(client, "remove", {"graph_name": "ex"})["graph_type"] == "undirected"
    # weighted edge list -> sssp
    p = tmp_path / "w.el"
    p.write_text("0 1 0.5\n1 2 1.25\n0 2 4.0\n2 3 1.0\n")
    act(client, "create", {"graph_name": "w", "file_format": "EdgeListWeighted", "path": str(p)})
    s = act(client, "compute", {"graph_name": "w", "property_key": "dist", "algorithm": {"Sssp": {"start_node": 0, "delta": 1.0}}})
    assert prop(client, s["property_id"]).column(0).to_pylist() == [0.0, 0.5, 1.75, 2.75]


def test_flight_put_then_page_rank_in_batches(client):
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
                                 "algorithm": {"PageRank": {"max_iterations": 20, "tolerance": 0.0, "damping_factor": 0.85}}})
    rd = client.do_get(fl.Ticket(json.dumps(pr["property_id"]).encode()))
    chunks = [c.data for c in rd]
    assert [c.num_rows for c in chunks[:-1]] == [10_000] * (len(chunks) - 1) and sum(c.num_rows for c in chunks) == n
    got = np.concatenate([c.column(0).to_numpy() for c in chunks])
    out = oracle.csr_build(src, dst, n, oracle.OUTGOING, oracle.SORTED)
    inc = oracle.csr_build(src, dst, n, oracle.INCOMING, oracle.SORTED)
    want, it, _ = oracle.page_rank_jacobi(inc[0], inc[1], out[0], 20, 0.0, 0.85, acc64=True)
    assert pr["algo_result"]["iterations"] == it == 20
    assert np.max(np.abs(got - want) / want) <= 1e-6
