"""The reference's own Python test-suite (crates/mate/tests/*.py) restated against graph_b200.
Same fixtures (conftest.py:5-30), same assertions, same order-dependence: `test_reorder` mutates the
module-scoped `ug` before the triangle-count golden 227874 is checked
(triangle_count_test.py:5-9 after graph_test.py:56-64).  Every call runs on the GPU through the
C ABI."""
import numpy as np
import pandas as pd
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gm():
    import graph_b200
    return graph_b200


@pytest.fixture(scope="module")
def g(gm, golden_dir):
    return gm.DiGraph.load(str(golden_dir / "scale_8.graph500"), layout=gm.Layout.Sorted)


@pytest.fixture(scope="module")
def ug(gm, golden_dir):
    return gm.Graph.load(str(golden_dir / "scale_8.graph500"), layout=gm.Layout.Sorted)


@pytest.fixture(scope="module")
def el_g(gm, golden_dir):
    return gm.DiGraph.load(str(golden_dir / "test.el"), layout=gm.Layout.Sorted, file_format=gm.FileFormat.EdgeList)


@pytest.fixture(scope="module")
def el_ug(gm, golden_dir):
    return gm.Graph.load(str(golden_dir / "test.el"), layout=gm.Layout.Sorted, file_format=gm.FileFormat.EdgeList)


# ---- ds_test.py ------------------------------------------------------------------------------
def test_numpy_graph(gm):
    el = np.array([[0, 1], [2, 3], [4, 1]], dtype=np.uint32)
    g = gm.Graph.from_numpy(el, layout=gm.Layout.Sorted)
    assert g.node_count() == 5
    assert g.edge_count() == 3
    assert np.array_equal(g.neighbors(0), np.array([1], dtype=np.uint32))
    assert np.array_equal(g.neighbors(1), np.array([0, 4], dtype=np.uint32))
    assert np.array_equal(g.neighbors(2), np.array([3], dtype=np.uint32))
    assert np.array_equal(g.neighbors(3), np.array([2], dtype=np.uint32))
    assert np.array_equal(g.neighbors(4), np.array([1], dtype=np.uint32))


def test_pandas_graph(gm):
    df = pd.DataFrame({"source": [0, 2, 4], "target": [1, 3, 1]})
    g = gm.Graph.from_pandas(df, layout=gm.Layout.Sorted)
    assert g.node_count() == 5
    assert g.edge_count() == 3
    assert np.array_equal(g.neighbors(1), np.array([0, 4], dtype=np.uint32))
    assert np.array_equal(g.neighbors(4), np.array([1], dtype=np.uint32))


def test_numpy_digraph(gm):
    el = np.array([[0, 1], [2, 3], [4, 1]], dtype=np.uint32)
    g = gm.DiGraph.from_numpy(el, layout=gm.Layout.Sorted)
    assert g.node_count() == 5
    assert g.edge_count() == 3
    assert np.array_equal(g.out_neighbors(0), np.array([1], dtype=np.uint32))
    assert np.array_equal(g.out_neighbors(2), np.array([3], dtype=np.uint32))
    assert np.array_equal(g.out_neighbors(4), np.array([1], dtype=np.uint32))
    assert np.array_equal(g.in_neighbors(1), np.array([0, 4], dtype=np.uint32))
    assert np.array_equal(g.in_neighbors(3), np.array([2], dtype=np.uint32))


def test_pandas_digraph(gm):
    df = pd.DataFrame({"source": [0, 2, 4], "target": [1, 3, 1]})
    g = gm.DiGraph.from_pandas(df, layout=gm.Layout.Sorted)
    assert g.node_count() == 5
    assert g.edge_count() == 3
    assert np.array_equal(g.in_neighbors(1), np.array([0, 4], dtype=np.uint32))


# ---- graph_edgelist_test.py --------------------------------------------------------------------
def test_load_edge_list_graph(el_g):
    assert el_g.node_count() == 5
    assert el_g.edge_count() == 6
    assert np.array_equal(el_g.out_neighbors(0), [1, 2])
    assert np.array_equal(el_g.out_neighbors(1), [2, 3])
    assert np.array_equal(el_g.out_neighbors(2), [4])
    assert np.array_equal(el_g.out_neighbors(3), [4])
    assert np.array_equal(el_g.out_neighbors(4), [])


def test_load_undirected_edge_list_graph(el_ug):
    assert el_ug.node_count() == 5
    assert el_ug.edge_count() == 6
    assert np.array_equal(el_ug.neighbors(0), [1, 2])
    assert np.array_equal(el_ug.neighbors(1), [0, 2, 3])
    assert np.array_equal(el_ug.neighbors(2), [0, 1, 4])
    assert np.array_equal(el_ug.neighbors(3), [1, 4])
    assert np.array_equal(el_ug.neighbors(4), [2, 3])


# ---- graph_test.py -----------------------------------------------------------------------------
def test_load_graph(g):
    assert g.node_count() == 1 << 8
    assert g.edge_count() == 1 << 12


def test_to_undirected(g, ug):
    undirected = g.to_undirected()
    for n in range(undirected.node_count()):
        assert set(undirected.copy_neighbors(n)) == set(ug.copy_neighbors(n))


def test_to_undirected_with_layout(gm):
    g = gm.DiGraph.from_numpy(np.array([[0, 1], [0, 1], [0, 2], [1, 2], [2, 1], [0, 3]], dtype=np.uint32))

    def compare_unsorted(expect, actual):
        s = expect.copy()
        s.sort()
        return np.array_equal(s, actual)

    for layout in (None, gm.Layout.Unsorted):
        ug = g.to_undirected(layout) if layout else g.to_undirected()
        assert compare_unsorted(ug.neighbors(0), [1, 1, 2, 3])
        assert compare_unsorted(ug.neighbors(1), [0, 0, 2, 2])
        assert compare_unsorted(ug.neighbors(2), [0, 1, 1])
        assert compare_unsorted(ug.neighbors(3), [0])
    ug = g.to_undirected(gm.Layout.Sorted)
    assert np.array_equal(ug.neighbors(0), [1, 1, 2, 3])
    assert np.array_equal(ug.neighbors(1), [0, 0, 2, 2])
    assert np.array_equal(ug.neighbors(2), [0, 1, 1])
    assert np.array_equal(ug.neighbors(3), [0])
    ug = g.to_undirected(gm.Layout.Deduplicated)
    assert np.array_equal(ug.neighbors(0), [1, 2, 3])
    assert np.array_equal(ug.neighbors(1), [0, 2])
    assert np.array_equal(ug.neighbors(2), [0, 1])
    assert np.array_equal(ug.neighbors(3), [0])


def test_reorder(ug):
    sorted_degrees = sorted((ug.degree(n) for n in range(ug.node_count())), reverse=True)
    ug.make_degree_ordered()
    degrees = [ug.degree(n) for n in range(ug.node_count())]
    assert degrees == sorted_degrees


def test_reorder_refused_while_views_alive(gm, golden_dir):
    h = gm.Graph.load(str(golden_dir / "scale_8.graph500"), layout=gm.Layout.Sorted)
    nb = h.neighbors(3)
    with pytest.raises(ValueError, match="cannot be reordered"):
        h.make_degree_ordered()  # crates/mate/src/graphs/mod.rs:264-276
    del nb
    h.make_degree_ordered()


# ---- numpy_neighbors_test.py -------------------------------------------------------------------
def test_out_neighbors(g):
    for n in range(g.node_count()):
        nb = g.out_neighbors(n)
        assert len(nb) == g.out_degree(n)
        assert nb.base is not None
        assert nb.tolist() == g.copy_out_neighbors(n)


def test_in_neighbors(g):
    for n in range(g.node_count()):
        nb = g.in_neighbors(n)
        assert len(nb) == g.in_degree(n)
        assert nb.base is not None
        assert nb.tolist() == g.copy_in_neighbors(n)


def test_neighbors(ug):
    for n in range(ug.node_count()):
        nb = ug.neighbors(n)
        assert len(nb) == ug.degree(n)
        assert nb.base is not None
        assert nb.tolist() == ug.copy_neighbors(n)


def test_neighbors_keep_alive(gm, golden_dir):
    g = gm.DiGraph.load(str(golden_dir / "scale_8.graph500"), layout=gm.Layout.Sorted)
    degree = g.in_degree(82)
    nb = g.in_neighbors(82)
    del g
    assert len(nb) == degree
    assert np.all([nb >= 0, nb < 1 << 8])
    with pytest.raises(ValueError):
        nb[0] = 1  # read-only view (shared_slice.rs:128)


# ---- page_rank_test.py -------------------------------------------------------------------------
def test_page_rank(g):
    pr = g.page_rank()
    assert pr.ran_iterations >= 1
    assert pr.error < 1.0
    assert pr.micros > 0
    scores = pr.scores()
    assert len(scores) == 1 << 8
    for score in scores:
        assert score > 0.0


def test_pr_max_iterations(g):
    assert g.page_rank(max_iterations=1).ran_iterations == 1


def test_pr_tolerance(g):
    assert g.page_rank(tolerance=1).ran_iterations == 1


def test_pr_damping_factor(g):
    pr = g.page_rank(damping_factor=0)
    assert pr.ran_iterations == 1
    for score in pr.scores():
        assert score == 1 / (1 << 8)


def test_pr_config_must_be_kwargs(g):
    with pytest.raises(TypeError):
        g.page_rank(42, 1.0, 0.1)


# ---- triangle_count_test.py --------------------------------------------------------------------
def test_triangle_count(ug):
    tc = ug.global_triangle_count()  # `ug` was degree-ordered by test_reorder above
    assert tc.triangles == 227874
    assert tc.micros > 0


@pytest.mark.parametrize("edges", [
    [[0, 1], [1, 2], [2, 0], [3, 4], [4, 5], [5, 3]],
    [[0, 1], [1, 2], [2, 0], [0, 3], [3, 4], [4, 0]],
    [[0, 1], [1, 2], [2, 0], [1, 3], [3, 2]],
])
def test_tc_small(gm, edges):
    ug = gm.Graph.from_numpy(np.array(edges, dtype=np.uint32), layout=gm.Layout.Deduplicated)
    assert ug.global_triangle_count().triangles == 2


# ---- wcc_test.py -------------------------------------------------------------------------------
def test_wcc(g):
    wcc = g.wcc()
    assert wcc.micros > 0
    components = wcc.components()
    assert len(components) == 1 << 8
    for component in components:
        assert component >= 0
        assert component < g.node_count()


def test_wcc_config_must_be_kwargs(g):
    with pytest.raises(TypeError):
        g.wcc(42, 1.0, 0.1)
