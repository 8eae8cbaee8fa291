// CPU-only check of GraphBuilder::file_format(..).path(..) (include/graph_b200.hpp): the reference's
// fixtures are read through the library's native readers; no device call is made.
//   crates/builder/tests/builder.rs:449-468 (scale_8.graph500), :493-564 (test.el)
#include <cstdio>

#include "graph_b200.hpp"

using namespace graph::prelude;

#define EXPECT(cond)                                                \
  do {                                                              \
    if (!(cond)) {                                                  \
      std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); \
      return 1;                                                     \
    }                                                               \
  } while (0)

int main(int argc, char** argv) {
  if (argc < 2) return 3;
  const std::string dir = argv[1];
  GraphBuilder b;
  b.file_format(FileFormat::Graph500).path(dir + "/scale_8.graph500");
  EXPECT(b.pending_edge_count() == 4096 && b.pending_node_count() == 256);
  std::size_t self_loops = 0;
  for (std::size_t i = 0; i < 4096; ++i) self_loops += b.pending_sources()[i] == b.pending_targets()[i];
  EXPECT(self_loops == 85);
  b.file_format(FileFormat::EdgeList).path(dir + "/test.el");
  EXPECT(b.pending_edge_count() == 6 && b.pending_node_count() == 0);
  const std::uint32_t s[6] = {0, 0, 1, 1, 2, 3}, t[6] = {1, 2, 2, 3, 4, 4};
  for (int i = 0; i < 6; ++i) EXPECT(b.pending_sources()[i] == s[i] && b.pending_targets()[i] == t[i]);
  b.path(dir + "/windows.el");  // CRLF line ends
  EXPECT(b.pending_edge_count() == 3 && b.pending_sources()[2] == 1 && b.pending_targets()[2] == 3);
  // weighted edge list: the value column is kept when asked for, dropped otherwise
  b.with_values().path(dir + "/test.wel");
  EXPECT(b.pending_edge_count() == 6 && b.pending_values().size() == 6);
  EXPECT(b.pending_values()[0] == 0.1f && b.pending_values()[5] == 0.6f);
  b.with_values(false).path(dir + "/test.wel");
  EXPECT(b.pending_edge_count() == 6 && b.pending_values().empty());
  bool threw = false;
  try { b.path(dir + "/does_not_exist.el"); } catch (const graph::Error&) { threw = true; }
  EXPECT(threw);
  std::printf("builder_io ok\n");
  return 0;
}
