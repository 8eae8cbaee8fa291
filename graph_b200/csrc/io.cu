// io.cu — native host-side readers of the reference's two input formats (no device code).
//
//   Graph500 packed edges   crates/builder/src/input/graph500.rs:63-127
//   text edge lists         crates/builder/src/input/edgelist.rs:181-279
//
// Like the reference the work is split into one contiguous chunk per hardware thread (edgelist.rs:
// 186-212 cuts at line boundaries); unlike it the chunks are written at prefix offsets, so the edge
// order is the file order on every run (the reference appends chunks in completion order).
#include <algorithm>
#include <charconv>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "common.cuh"

namespace gb {

static unsigned io_threads(uint64_t work_items) {
  unsigned hw = std::thread::hardware_concurrency();
  if (hw == 0) hw = 4;  // DEFAULT_PARALLELISM, crates/algos/src/lib.rs:152
  const uint64_t by_work = std::max<uint64_t>(1, work_items / (1u << 16));
  return (unsigned)std::min<uint64_t>(hw, by_work);
}

template <typename F>
static void parallel_chunks(unsigned threads, F&& body) {
  if (threads <= 1) {
    body(0u);
    return;
  }
  std::vector<std::thread> pool;
  pool.reserve(threads);
  for (unsigned t = 0; t < threads; ++t) pool.emplace_back([&body, t] { body(t); });
  for (auto& th : pool) th.join();
}

// one line "<src><1 byte><dst>[ <value>]<newline>"; returns the position after the line.
// Nothing is read at or beyond the line's end (the buffer is a (pointer, length) pair, not a C string):
// the value is parsed by std::from_chars over [p, end of line) — bounded, locale-free, and like the
// reference's fast_float2::parse_partial (edgelist.rs:237-241) it takes the longest valid prefix.
static inline uint64_t parse_line(const char* text, uint64_t p, uint64_t len, uint64_t nl, uint64_t* s, uint64_t* t,
                                  float* v) {
  (void)nl;
  const void* nlp = std::memchr(text + p, '\n', len - p);
  const uint64_t eol = nlp ? (uint64_t)(static_cast<const char*>(nlp) - text) : len;  // position of '\n' (or len)
  uint64_t a = 0, b = 0;
  while (p < eol && text[p] >= '0' && text[p] <= '9') a = a * 10 + (uint64_t)(text[p++] - '0');
  if (p < eol) p += 1;  // exactly one separator byte (edgelist.rs:225)
  while (p < eol && text[p] >= '0' && text[p] <= '9') b = b * 10 + (uint64_t)(text[p++] - '0');
  float val = 0.0f;  // EV::default() when the column is missing (edgelist.rs:237-241)
  if (p < eol && text[p] == ' ') {
    ++p;
    if (p < eol && text[p] == '+') ++p;
    float parsed = 0.0f;
    const auto res = std::from_chars(text + p, text + eol, parsed);
    if (res.ec == std::errc() || res.ec == std::errc::result_out_of_range) val = parsed;
  }
  *s = a;
  *t = b;
  *v = val;
  return eol < len ? eol + 1 : len;
}

}  // namespace gb

extern "C" {

// PackedEdge{v0_low, v1_low, high}: src = v0_low | (high & 0xFFFF) << 32, dst = v1_low | (high >> 16) << 32
// (graph500.rs:111-127); node_count = edge_count / 16 (graph500.rs:74).
gb_status gb_graph500_decode(const void* bytes, uint64_t len, uint32_t* src, uint32_t* dst, uint64_t* edge_count,
                             uint32_t* node_count) {
  GB_REQUIRE(edge_count && node_count, "NULL argument");
  const uint64_t m = len / 12;
  *edge_count = m;
  *node_count = (uint32_t)std::min<uint64_t>(m / 16, 0xFFFFFFFFull);
  if (m == 0) return GB_OK;
  GB_REQUIRE(bytes && src && dst, "NULL argument");
  const unsigned T = gb::io_threads(m);
  std::vector<int> bad(T, 0);
  const uint8_t* base = static_cast<const uint8_t*>(bytes);
  gb::parallel_chunks(T, [&](unsigned t) {
    const uint64_t b = m * t / T, e = m * (t + 1) / T;
    for (uint64_t i = b; i < e; ++i) {
      uint32_t rec[3];
      std::memcpy(rec, base + 12 * i, 12);
      if (rec[2] != 0) bad[t] = 1;  // an id above 32 bits: Idx::new asserts (index.rs:51-54)
      src[i] = rec[0];
      dst[i] = rec[1];
    }
  });
  for (int b : bad) GB_REQUIRE(b == 0, "Graph500 node id does not fit 32 bits");
  return GB_OK;
}

// The inverse: edges -> packed 12-byte records (high word 0 for 32-bit ids), the file format the
// reference's CLI reads with `-f graph500 --use-32-bit` (crates/app/src/runner.rs:104-133).
gb_status gb_graph500_encode(const uint32_t* src, const uint32_t* dst, uint64_t edge_count, void* bytes) {
  if (edge_count == 0) return GB_OK;
  GB_REQUIRE(src && dst && bytes, "NULL argument");
  const unsigned T = gb::io_threads(edge_count);
  uint8_t* base = static_cast<uint8_t*>(bytes);
  gb::parallel_chunks(T, [&](unsigned t) {
    const uint64_t b = edge_count * t / T, e = edge_count * (t + 1) / T;
    for (uint64_t i = b; i < e; ++i) {
      const uint32_t rec[3] = {src[i], dst[i], 0u};
      std::memcpy(base + 12 * i, rec, 12);
    }
  });
  return GB_OK;
}

// Two-phase use: call with src == NULL to obtain *edge_count, allocate, call again to fill.
// values may be NULL.  Ids above 32 bits are an error.
gb_status gb_edge_list_parse(const char* text, uint64_t len, uint32_t* src, uint32_t* dst, float* values,
                             uint64_t* edge_count) {
  GB_REQUIRE(edge_count, "NULL argument");
  *edge_count = 0;
  if (len == 0) return GB_OK;
  GB_REQUIRE(text != nullptr, "NULL text");
  // new_line_bytes, edgelist.rs:271-279
  uint64_t nl = 1;
  if (const void* first = std::memchr(text, '\n', len)) {
    const uint64_t i = (uint64_t)(static_cast<const char*>(first) - text);
    if (i > 0 && text[i - 1] == '\r') nl = 2;
  }
  // chunk boundaries moved forward to the next line start (edgelist.rs:193-212)
  const unsigned T = gb::io_threads(len / 8);
  std::vector<uint64_t> start(T + 1, len);
  start[0] = 0;
  for (unsigned t = 1; t < T; ++t) {
    uint64_t p = len * t / T;
    while (p < len && text[p - 1] != '\n') ++p;
    start[t] = std::max(p, start[t - 1]);
  }
  std::vector<uint64_t> count(T + 1, 0);
  std::vector<int> bad(T, 0);
  gb::parallel_chunks(T, [&](unsigned t) {
    uint64_t c = 0;
    for (uint64_t p = start[t]; p < start[t + 1];) {
      uint64_t s, d;
      float v;
      p = gb::parse_line(text, p, len, nl, &s, &d, &v);
      if (s > 0xFFFFFFFFull || d > 0xFFFFFFFFull) bad[t] = 1;
      ++c;
    }
    count[t + 1] = c;
  });
  for (int b : bad) GB_REQUIRE(b == 0, "edge list node id does not fit 32 bits");
  for (unsigned t = 0; t < T; ++t) count[t + 1] += count[t];
  *edge_count = count[T];
  if (!src) return GB_OK;
  GB_REQUIRE(dst != nullptr, "dst is NULL");
  gb::parallel_chunks(T, [&](unsigned t) {
    uint64_t i = count[t];
    for (uint64_t p = start[t]; p < start[t + 1]; ++i) {
      uint64_t s, d;
      float v;
      p = gb::parse_line(text, p, len, nl, &s, &d, &v);
      src[i] = (uint32_t)s;
      dst[i] = (uint32_t)d;
      if (values) values[i] = v;
    }
  });
  return GB_OK;
}

}  // extern "C"
