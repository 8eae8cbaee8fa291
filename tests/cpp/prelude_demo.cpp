// The reference's doc-tests / unit tests for the hot path, written against include/graph_b200.hpp
// (graph::prelude mirror).  Exits 0 when every golden matches; prints the first mismatch otherwise.
//   crates/algos/src/lib.rs:96-140 (page_rank, 13 nodes), page_rank.rs:176-197, sssp.rs:283-313,
//   triangle_count.rs:94-130, wcc.rs:307-329, graph_ops.rs:718-774
#include <cstdio>
#include <cstring>

#include "graph_b200.hpp"

using namespace graph::prelude;

#define EXPECT(cond)                                              \
  do {                                                            \
    if (!(cond)) {                                                \
      std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); \
      return 1;                                                   \
    }                                                             \
  } while (0)

int main(int argc, char** argv) {
  try {
    {  // page_rank doc-test, lib.rs:96-140
      DirectedCsrGraph g = GraphBuilder()
                               .edges({{1, 2}, {2, 1}, {4, 0}, {4, 1}, {5, 4}, {5, 1}, {5, 6}, {6, 1}, {6, 5}, {7, 1},
                                       {7, 5}, {8, 1}, {8, 5}, {9, 1}, {9, 5}, {10, 1}, {10, 5}, {11, 5}, {12, 5}})
                               .build_directed();
      auto [ranks, iterations, error] = page_rank(g, PageRankConfig::new_(10, 1e-4, 0.85f));
      (void)error;
      const float expected[13] = {0.024064068f, 0.3145448f,  0.27890152f, 0.01153846f, 0.029471997f, 0.06329483f, 0.029471997f,
                                  0.01153846f,  0.01153846f, 0.01153846f, 0.01153846f, 0.01153846f,  0.01153846f};
      EXPECT(iterations == 10);
      EXPECT(ranks.size() == 13 && std::memcmp(ranks.data(), expected, sizeof expected) == 0);
    }
    {  // test_pr_two_components, page_rank.rs:176-197
      DirectedCsrGraph g = GraphBuilder().csr_layout(CsrLayout::Sorted).edges({{0, 1}, {1, 2}, {0, 2}, {3, 4}, {4, 5}, {3, 5}}).build_directed();
      auto [ranks, it, err] = page_rank(g, PageRankConfig{});
      (void)it; (void)err;
      const float expected[6] = {0.024999997f, 0.035624996f, 0.06590624f, 0.024999997f, 0.035624996f, 0.06590624f};
      EXPECT(std::memcmp(ranks.data(), expected, sizeof expected) == 0);
    }
    {  // test_sssp, sssp.rs:283-313
      DirectedCsrGraph g = GraphBuilder()
                               .csr_layout(CsrLayout::Deduplicated)
                               .edges_with_values({{0, 1, 4.f}, {0, 2, 2.f}, {1, 2, 5.f}, {1, 3, 10.f}, {2, 4, 3.f}, {3, 5, 11.f}, {4, 3, 4.f}})
                               .build_directed();
      std::vector<float> d = delta_stepping(g, DeltaSteppingConfig::new_(0, 3.0f));
      const float expected[6] = {0.f, 4.f, 2.f, 9.f, 5.f, 20.f};
      EXPECT(d.size() == 6 && std::memcmp(d.data(), expected, sizeof expected) == 0);
    }
    {  // triangle_count.rs:94-130
      auto tc = [](std::vector<std::pair<std::uint32_t, std::uint32_t>> e) {
        return global_triangle_count(GraphBuilder().csr_layout(CsrLayout::Deduplicated).edges(e).build_undirected());
      };
      EXPECT(tc({{0, 1}, {1, 2}, {0, 2}, {3, 4}, {4, 5}, {3, 5}}) == 2);
      EXPECT(tc({{0, 1}, {1, 2}, {0, 2}, {0, 3}, {3, 4}, {0, 4}}) == 2);
      EXPECT(tc({{0, 1}, {1, 2}, {0, 2}, {1, 3}, {2, 3}}) == 2);
    }
    {  // two_components_afforest, wcc.rs:318-329
      DirectedCsrGraph g = GraphBuilder().edges({{0, 1}, {2, 3}}).build_directed();
      Components res = wcc_afforest(g, WccConfig{});
      EXPECT(res.component(0) == res.component(1));
      EXPECT(res.component(2) == res.component(3));
      EXPECT(res.component(1) != res.component(2));
    }
    {  // relabel_by_degree_test, graph_ops.rs:741-774
      UndirectedCsrGraph g = GraphBuilder().edges({{0, 1}, {1, 2}, {1, 3}, {2, 0}, {2, 1}, {2, 3}, {3, 0}, {3, 2}}).build_undirected();
      relabel_graph(g);
      EXPECT(g.degree(0) == 5 && g.degree(1) == 4 && g.degree(2) == 4 && g.degree(3) == 3);
      const std::uint32_t n0[5] = {1, 1, 2, 2, 3};
      auto r = g.neighbors(0);
      EXPECT(r.second - r.first == 5 && std::memcmp(r.first, n0, sizeof n0) == 0);
    }
    if (argc > 1) {  // builder.rs doc-tests: file inputs (argv[1] = directory with the reference's fixtures)
      const std::string dir = argv[1];
      DirectedCsrGraph g = GraphBuilder().csr_layout(CsrLayout::Sorted).file_format(FileFormat::Graph500)
                               .path(dir + "/scale_8.graph500").build_directed();
      EXPECT(g.node_count() == 256 && g.edge_count() == 4096);   // crates/builder/tests/builder.rs:449-468
      auto o = g.out_neighbors(0);
      EXPECT(o.second - o.first == 2 && o.first[0] == 37 && o.first[1] == 157);
      UndirectedCsrGraph u = GraphBuilder().csr_layout(CsrLayout::Sorted).file_format(FileFormat::EdgeList)
                                 .path(dir + "/test.el").build_undirected();
      EXPECT(u.node_count() == 5 && u.edge_count() == 6 && u.degree(1) == 3);   // builder.rs:534-564
      relabel_graph(u);
    }
    {  // errors instead of panics
      DirectedCsrGraph g = GraphBuilder().edges({{0, 1}}).build_directed();
      bool threw = false;
      try { delta_stepping(g, DeltaSteppingConfig::new_(0, 1.0f)); } catch (const graph::Error&) { threw = true; }
      EXPECT(threw);  // no edge values: DirectedNeighborsWithValues<NI, f32> is not implemented for EV = ()
    }
  } catch (const graph::Error& e) {
    std::printf("graph::Error %d: %s\n", (int)e.status, e.what());
    return 2;
  }
  std::printf("prelude_demo ok\n");
  return 0;
}
