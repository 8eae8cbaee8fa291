// graph.cu — device-resident twin of the reference's CSR graph types and its construction.
//
//   DirectedCsrGraph{csr_out, csr_inc}   crates/builder/src/graph/csr.rs:364-368
//   UndirectedCsrGraph{csr}              crates/builder/src/graph/csr.rs:658-661
//   Csr::from((&edges, n, direction, layout))   csr.rs:124-221  -> build_csr_device (radix sort)
//   to_undirected                        csr.rs:391-464
//   make_degree_ordered                  crates/builder/src/graph_ops.rs:511-638
//
// The reference builds a CSR with an atomic scatter followed by a per-row sort; here the whole
// build is one device radix sort of packed (row, target) keys, which yields the Sorted layout
// directly, makes Unsorted deterministic (stable sort on the row bits only = edge-list order) and
// turns Deduplicated into a flagged compaction.
#include <cub/cub.cuh>

#include "common.cuh"
#include "rmat.cuh"

namespace gb {

std::string& last_error() {
  static thread_local std::string msg;
  return msg;
}

gb_status fail(gb_status st, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  last_error() = buf;
  return st;
}

static bool g_profiling = false;
bool profiling_on() { return g_profiling; }

static inline uint32_t bits_for(uint32_t n) {
  uint32_t b = 1;
  while (b < 32 && (1ull << b) < n) ++b;
  return b;
}

// ---- kernels -------------------------------------------------------------------------------
__global__ void k_pack_keys(const uint32_t* __restrict__ rows, const uint32_t* __restrict__ cols,
                            uint64_t count, uint32_t bits, uint64_t* __restrict__ keys) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < count;
       i += (uint64_t)gridDim.x * blockDim.x)
    keys[i] = ((uint64_t)rows[i] << bits) | cols[i];
}

__global__ void k_iota(uint32_t* __restrict__ a, uint64_t count) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < count;
       i += (uint64_t)gridDim.x * blockDim.x)
    a[i] = (uint32_t)i;
}

// offsets from a row-sorted sequence: the end of each non-empty row's run is marked at
// marks[row + 1]; offsets are the running maximum of the marks (empty rows inherit the previous
// end), so no thread ever walks a gap of empty rows.
template <typename RowFn>
__global__ void k_mark_row_ends(RowFn row_of, uint64_t count, uint32_t* __restrict__ marks) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < count;
       i += (uint64_t)gridDim.x * blockDim.x) {
    uint32_t r = row_of(i);
    if (i + 1 == count || row_of(i + 1) != r) marks[r + 1] = (uint32_t)(i + 1);
  }
}

struct RowOfKey {
  const uint64_t* keys;
  uint32_t bits;
  __device__ uint32_t operator()(uint64_t i) const { return (uint32_t)(keys[i] >> bits); }
};
struct RowOfArr {
  const uint32_t* rows;
  __device__ uint32_t operator()(uint64_t i) const { return rows[i]; }
};

template <typename RowFn>
static gb_status offsets_from_sorted(cudaStream_t s, RowFn row_of, uint64_t count, uint32_t n,
                                     uint32_t* off) {
  GB_CUDA(cudaMemsetAsync(off, 0, ((size_t)n + 1) * 4, s));
  if (count) k_mark_row_ends<<<grid_for(count, 256), 256, 0, s>>>(row_of, count, off);
  size_t tb = 0;
  GB_CUDA(cub::DeviceScan::InclusiveScan(nullptr, tb, off, off, cub::Max(), (int64_t)n + 1, s));
  DevBuf<uint8_t> tmp;
  GB_TRY(tmp.alloc(tb));
  GB_CUDA(cub::DeviceScan::InclusiveScan(tmp.p, tb, off, off, cub::Max(), (int64_t)n + 1, s));
  GB_CUDA(cudaStreamSynchronize(s));  // tmp is released on return
  return GB_OK;
}

__global__ void k_unpack_targets(const uint64_t* __restrict__ keys, uint64_t count, uint32_t bits,
                                 uint32_t* __restrict__ tgt) {
  uint64_t mask = (bits >= 32) ? 0xFFFFFFFFull : ((1ull << bits) - 1ull);
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < count;
       i += (uint64_t)gridDim.x * blockDim.x)
    tgt[i] = (uint32_t)(keys[i] & mask);
}

template <typename T>
__global__ void k_gather(const T* __restrict__ src, const uint32_t* __restrict__ idx, uint64_t count,
                         T* __restrict__ dst) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < count;
       i += (uint64_t)gridDim.x * blockDim.x)
    dst[i] = src[idx[i]];
}

// Deduplicated layout (csr.rs:897-948): keep the first of each run of equal (row,target) keys and
// drop entries whose target is the row itself.
__global__ void k_dedup_flags(const uint64_t* __restrict__ keys, uint64_t count, uint32_t bits,
                              uint8_t* __restrict__ flags) {
  uint64_t mask = (bits >= 32) ? 0xFFFFFFFFull : ((1ull << bits) - 1ull);
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < count;
       i += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t k = keys[i];
    bool keep = (i == 0 || keys[i - 1] != k) && ((uint32_t)(k >> bits) != (uint32_t)(k & mask));
    flags[i] = keep ? 1 : 0;
  }
}

// expands CSR offsets into one row id per entry: one warp per row (coalesced stores, no searches)
__global__ void k_expand_rows(const uint32_t* __restrict__ off, uint32_t n, uint64_t count,
                              uint32_t* __restrict__ rows) {
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t v = warp; v < n; v += nwarps) {
    const uint32_t b = off[v], e = off[v + 1];
    for (uint32_t i = b + lane; i < e; i += 32) rows[i] = v;
  }
  (void)count;
}

__global__ void k_check_ids(const uint32_t* __restrict__ a, uint64_t count, uint32_t n,
                            unsigned int* __restrict__ bad) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < count;
       i += (uint64_t)gridDim.x * blockDim.x)
    if (a[i] >= n) atomicAdd(bad, 1u);
}

__global__ void k_rmat(uint32_t scale, uint64_t seed, uint64_t first, uint64_t count,
                       uint32_t* __restrict__ src, uint32_t* __restrict__ dst) {
  RmatScramble scr(scale, seed);
  for (uint64_t k = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; k < count;
       k += (uint64_t)gridDim.x * blockDim.x) {
    uint32_t s, t;
    rmat_edge(scale, seed, first + k, scr, &s, &t);
    src[k] = s;
    dst[k] = t;
  }
}

__global__ void k_rmat_weights(uint64_t seed, uint64_t count, float* __restrict__ w) {
  for (uint64_t k = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; k < count;
       k += (uint64_t)gridDim.x * blockDim.x)
    w[k] = rmat_weight(seed, k);
}

// make_degree_ordered helpers
__global__ void k_degree_keys(const uint32_t* __restrict__ off, uint32_t n, uint64_t* __restrict__ keys) {
  for (uint32_t v = blockIdx.x * blockDim.x + threadIdx.x; v < n; v += gridDim.x * blockDim.x)
    keys[v] = ((uint64_t)(off[v + 1] - off[v]) << 32) | v;
}
__global__ void k_rank_to_newid(const uint64_t* __restrict__ sorted, uint32_t n,
                                uint32_t* __restrict__ new_id, uint32_t* __restrict__ new_deg) {
  for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
    uint64_t k = sorted[r];
    new_id[(uint32_t)k] = r;                 // unzip_degrees_and_nodes, graph_ops.rs:564-592
    new_deg[r] = (uint32_t)(k >> 32);
  }
}
__global__ void k_relabel_keys(const uint32_t* __restrict__ rows, const uint32_t* __restrict__ tgt,
                               const uint32_t* __restrict__ new_id, uint64_t count, uint32_t bits,
                               uint64_t* __restrict__ keys) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < count;
       i += (uint64_t)gridDim.x * blockDim.x)
    keys[i] = ((uint64_t)new_id[rows[i]] << bits) | new_id[tgt[i]];
}

// ---- CSR build -----------------------------------------------------------------------------
gb_status build_csr_device(cudaStream_t s, uint32_t n, uint32_t* d_rows, uint32_t* d_cols, float* d_w,
                           uint64_t count, gb_layout layout, DevCsr* csr) {
  GB_REQUIRE(count < 0xFFFFFFFFull, "CSR with %llu entries does not fit u32 offsets (csr.rs:124)",
             (unsigned long long)count);
  const uint32_t bits = bits_for(n);
  const unsigned blk = 256;
  GB_TRY(csr->off.alloc((size_t)n + 1));
  csr->len = count;
  if (count == 0) {
    GB_CUDA(cudaMemsetAsync(csr->off.p, 0, ((size_t)n + 1) * 4, s));
    GB_TRY(csr->tgt.alloc(0, 8));
    GB_CUDA(cudaMemsetAsync(csr->tgt.p, 0, 8 * 4, s));
    if (d_w) GB_TRY(csr->w.alloc(0, 8));
    return GB_OK;
  }

  if (layout == GB_LAYOUT_UNSORTED) {
    // stable sort on the row id only: within a row the edge-list order survives
    DevBuf<uint32_t> keys_alt, idx, idx_alt, rows_copy;
    GB_TRY(rows_copy.alloc(count));
    GB_TRY(keys_alt.alloc(count));
    GB_TRY(idx.alloc(count));
    GB_TRY(idx_alt.alloc(count));
    GB_CUDA(cudaMemcpyAsync(rows_copy.p, d_rows, count * 4, cudaMemcpyDeviceToDevice, s));
    k_iota<<<grid_for(count, blk), blk, 0, s>>>(idx.p, count);
    cub::DoubleBuffer<uint32_t> kb(rows_copy.p, keys_alt.p);
    cub::DoubleBuffer<uint32_t> vb(idx.p, idx_alt.p);
    size_t tmp_bytes = 0;
    GB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, kb, vb, count, 0, (int)bits, s));
    DevBuf<uint8_t> tmp;
    GB_TRY(tmp.alloc(tmp_bytes));
    GB_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, kb, vb, count, 0, (int)bits, s));
    GB_TRY(csr->tgt.alloc(count, 8));
    GB_CUDA(cudaMemsetAsync(csr->tgt.p + count, 0, 8 * 4, s));
    k_gather<uint32_t><<<grid_for(count, blk), blk, 0, s>>>(d_cols, vb.Current(), count, csr->tgt.p);
    if (d_w) {
      GB_TRY(csr->w.alloc(count, 8));
      k_gather<float><<<grid_for(count, blk), blk, 0, s>>>(d_w, vb.Current(), count, csr->w.p);
    }
    GB_TRY(offsets_from_sorted(s, RowOfArr{kb.Current()}, count, n, csr->off.p));
    GB_CUDA(cudaGetLastError());
    GB_CUDA(cudaStreamSynchronize(s));
    return GB_OK;
  }

  // Sorted / Deduplicated: one radix sort of (row << bits | target) keys
  DevBuf<uint64_t> keys, keys_alt;
  GB_TRY(keys.alloc(count));
  GB_TRY(keys_alt.alloc(count));
  k_pack_keys<<<grid_for(count, blk), blk, 0, s>>>(d_rows, d_cols, count, bits, keys.p);
  cub::DoubleBuffer<uint64_t> kb(keys.p, keys_alt.p);
  DevBuf<uint32_t> idx, idx_alt;
  DevBuf<uint8_t> tmp;
  size_t tmp_bytes = 0;
  const uint32_t* order = nullptr;
  if (d_w) {
    GB_TRY(idx.alloc(count));
    GB_TRY(idx_alt.alloc(count));
    k_iota<<<grid_for(count, blk), blk, 0, s>>>(idx.p, count);
    cub::DoubleBuffer<uint32_t> vb(idx.p, idx_alt.p);
    GB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, kb, vb, count, 0, (int)(2 * bits), s));
    GB_TRY(tmp.alloc(tmp_bytes));
    GB_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, kb, vb, count, 0, (int)(2 * bits), s));
    order = vb.Current();
  } else {
    GB_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, kb, count, 0, (int)(2 * bits), s));
    GB_TRY(tmp.alloc(tmp_bytes));
    GB_CUDA(cub::DeviceRadixSort::SortKeys(tmp.p, tmp_bytes, kb, count, 0, (int)(2 * bits), s));
  }
  uint64_t* sorted = kb.Current();
  uint64_t* spare = kb.Alternate();
  uint64_t out_count = count;

  DevBuf<uint32_t> order_c;
  if (layout == GB_LAYOUT_DEDUPLICATED) {
    DevBuf<uint8_t> flags;
    DevBuf<uint64_t> d_num;
    GB_TRY(flags.alloc(count));
    GB_TRY(d_num.alloc(1));
    k_dedup_flags<<<grid_for(count, blk), blk, 0, s>>>(sorted, count, bits, flags.p);
    size_t sel_bytes = 0;
    GB_CUDA(cub::DeviceSelect::Flagged(nullptr, sel_bytes, sorted, flags.p, spare, d_num.p,
                                       (int64_t)count, s));
    DevBuf<uint8_t> sel_tmp;
    GB_TRY(sel_tmp.alloc(sel_bytes));
    GB_CUDA(cub::DeviceSelect::Flagged(sel_tmp.p, sel_bytes, sorted, flags.p, spare, d_num.p,
                                       (int64_t)count, s));
    if (order) {
      GB_TRY(order_c.alloc(count));
      // the uint32 selection has its own temporary-storage size: query it (never reuse the uint64 one)
      size_t sel_bytes32 = 0;
      GB_CUDA(cub::DeviceSelect::Flagged(nullptr, sel_bytes32, order, flags.p, order_c.p, d_num.p,
                                         (int64_t)count, s));
      DevBuf<uint8_t> sel_tmp32;
      GB_TRY(sel_tmp32.alloc(sel_bytes32));
      GB_CUDA(cub::DeviceSelect::Flagged(sel_tmp32.p, sel_bytes32, order, flags.p, order_c.p, d_num.p,
                                         (int64_t)count, s));
      GB_CUDA(cudaStreamSynchronize(s));  // sel_tmp32 is released at the end of this scope
      order = order_c.p;
    }
    GB_CUDA(cudaMemcpyAsync(&out_count, d_num.p, 8, cudaMemcpyDeviceToHost, s));
    GB_CUDA(cudaStreamSynchronize(s));
    sorted = spare;
  }

  csr->len = out_count;
  GB_TRY(csr->tgt.alloc(out_count, 8));
  GB_CUDA(cudaMemsetAsync(csr->tgt.p + out_count, 0, 8 * 4, s));
  if (out_count) {
    k_unpack_targets<<<grid_for(out_count, blk), blk, 0, s>>>(sorted, out_count, bits, csr->tgt.p);
    if (d_w) {
      GB_TRY(csr->w.alloc(out_count, 8));
      k_gather<float><<<grid_for(out_count, blk), blk, 0, s>>>(d_w, order, out_count, csr->w.p);
    }
  } else if (d_w) {
    GB_TRY(csr->w.alloc(0, 8));
  }
  GB_TRY(offsets_from_sorted(s, RowOfKey{sorted, bits}, out_count, n, csr->off.p));
  GB_CUDA(cudaGetLastError());
  GB_CUDA(cudaStreamSynchronize(s));
  return GB_OK;
}

gb_status new_graph(int device, gb_graph_kind kind, uint32_t n, gb_graph** out) {
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0)
    return fail(GB_ERR_CUDA, "no CUDA device available: libgraph_b200 has no CPU fallback");
  GB_REQUIRE(device >= 0 && device < count, "device %d out of range (have %d)", device, count);
  GB_CUDA(cudaSetDevice(device));
  gb_graph* g = new (std::nothrow) gb_graph();
  if (!g) return fail(GB_ERR_OOM, "host allocation failed");
  g->device = device;
  g->kind = kind;
  g->n = n;
  cudaError_t e = cudaStreamCreateWithFlags(&g->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreate(&g->ev_begin);
  if (e == cudaSuccess) e = cudaEventCreate(&g->ev_end);
  if (e != cudaSuccess) {
    delete g;
    return fail(GB_ERR_CUDA, "stream/event creation failed: %s", cudaGetErrorString(e));
  }
  *out = g;
  return GB_OK;
}

static gb_status upload_csr(cudaStream_t s, uint32_t n, const uint32_t* off, const uint32_t* tgt,
                            const float* w, DevCsr* csr) {
  uint64_t len = off[n];
  csr->len = len;
  GB_TRY(csr->off.alloc((size_t)n + 1));
  GB_TRY(csr->tgt.alloc(len, 8));
  GB_CUDA(cudaMemcpyAsync(csr->off.p, off, ((size_t)n + 1) * 4, cudaMemcpyHostToDevice, s));
  if (len) GB_CUDA(cudaMemcpyAsync(csr->tgt.p, tgt, len * 4, cudaMemcpyHostToDevice, s));
  GB_CUDA(cudaMemsetAsync(csr->tgt.p + len, 0, 8 * 4, s));
  if (w) {
    GB_TRY(csr->w.alloc(len, 8));
    if (len) GB_CUDA(cudaMemcpyAsync(csr->w.p, w, len * 4, cudaMemcpyHostToDevice, s));
  }
  return GB_OK;
}

// Host-side sanity of the offsets (O(n)); the O(m) target range check runs on the device after
// the upload (k_check_ids) so that a billion-edge twin is not validated by one CPU thread.
static gb_status validate_host_csr(uint32_t n, const uint32_t* off, const uint32_t* tgt, const char* what) {
  GB_REQUIRE(off != nullptr, "%s offsets is NULL", what);
  GB_REQUIRE(off[0] == 0, "%s offsets[0] must be 0", what);
  GB_REQUIRE(off[n] == 0 || tgt != nullptr, "%s targets is NULL", what);
  return GB_OK;
}

__global__ void k_check_monotone(const uint32_t* __restrict__ off, uint32_t n, unsigned int* __restrict__ bad) {
  for (uint32_t v = blockIdx.x * blockDim.x + threadIdx.x; v < n; v += gridDim.x * blockDim.x)
    if (off[v] > off[v + 1]) atomicAdd(bad, 1u);
}

static gb_status validate_device_targets(cudaStream_t s, uint32_t n, const DevCsr& c, const char* what) {
  DevBuf<unsigned int> bad;
  GB_TRY(bad.alloc(2));
  GB_CUDA(cudaMemsetAsync(bad.p, 0, 8, s));
  if (c.len) k_check_ids<<<grid_for(c.len, 256), 256, 0, s>>>(c.tgt.p, c.len, n, bad.p);
  k_check_monotone<<<grid_for(n, 256), 256, 0, s>>>(c.off.p, n, bad.p + 1);
  unsigned int nbad[2] = {0, 0};
  GB_CUDA(cudaMemcpyAsync(nbad, bad.p, 8, cudaMemcpyDeviceToHost, s));
  GB_CUDA(cudaStreamSynchronize(s));
  GB_REQUIRE(nbad[1] == 0, "%s offsets are not monotone (%u rows)", what, nbad[1]);
  GB_REQUIRE(nbad[0] == 0, "%s CSR holds %u targets >= node_count %u", what, nbad[0], n);
  return GB_OK;
}

gb_status upload_host_csr(cudaStream_t s, uint32_t n, const uint32_t* off, const uint32_t* tgt, const float* w,
                          DevCsr* csr, const char* what) {
  GB_TRY(validate_host_csr(n, off, tgt ? tgt : off, what));
  if (tgt) {
    GB_TRY(upload_csr(s, n, off, tgt, w, csr));
    return validate_device_targets(s, n, *csr, what);
  }
  // offsets only (degrees): no target array on the device
  csr->len = off[n];
  GB_TRY(csr->off.alloc((size_t)n + 1));
  GB_CUDA(cudaMemcpyAsync(csr->off.p, off, ((size_t)n + 1) * 4, cudaMemcpyHostToDevice, s));
  DevCsr none;
  (void)none;
  return GB_OK;
}

// builds a directed or undirected graph from DEVICE edge arrays
static gb_status graph_from_device_edges(gb_graph* g, uint32_t* d_src, uint32_t* d_dst, float* d_w,
                                         uint64_t m, gb_layout layout) {
  cudaStream_t s = g->stream;
  if (g->kind == GB_KIND_DIRECTED) {
    GB_TRY(build_csr_device(s, g->n, d_src, d_dst, d_w, m, layout, &g->out));
    GB_TRY(build_csr_device(s, g->n, d_dst, d_src, nullptr, m, layout, &g->in));
  } else {
    // one CSR holding both directions: outgoing pass first, then incoming (csr.rs:154-172)
    DevBuf<uint32_t> rows, cols;
    GB_TRY(rows.alloc(2 * m));
    GB_TRY(cols.alloc(2 * m));
    if (m) {
      GB_CUDA(cudaMemcpyAsync(rows.p, d_src, m * 4, cudaMemcpyDeviceToDevice, s));
      GB_CUDA(cudaMemcpyAsync(rows.p + m, d_dst, m * 4, cudaMemcpyDeviceToDevice, s));
      GB_CUDA(cudaMemcpyAsync(cols.p, d_dst, m * 4, cudaMemcpyDeviceToDevice, s));
      GB_CUDA(cudaMemcpyAsync(cols.p + m, d_src, m * 4, cudaMemcpyDeviceToDevice, s));
    }
    GB_TRY(build_csr_device(s, g->n, rows.p, cols.p, nullptr, 2 * m, layout, &g->out));
  }
  return GB_OK;
}

static gb_status graph_from_host_edges(int device, gb_graph_kind kind, const uint32_t* src,
                                       const uint32_t* dst, const float* w, uint64_t m, uint32_t n,
                                       gb_layout layout, gb_graph** out) {
  GB_REQUIRE(out != nullptr, "graph out-pointer is NULL");
  GB_REQUIRE(m == 0 || (src && dst), "edge arrays are NULL");
  GB_REQUIRE((int)layout >= 0 && (int)layout <= 2, "bad layout %d", (int)layout);
  uint64_t cap = (kind == GB_KIND_UNDIRECTED) ? 2 * m : m;
  GB_REQUIRE(cap < 0xFFFFFFFFull, "edge count %llu does not fit u32 offsets", (unsigned long long)m);
  if (n == 0) {  // Edges::max_node_id + 1, edgelist.rs:84-90
    uint32_t mx = 0;
    for (uint64_t i = 0; i < m; ++i) {
      if (src[i] > mx) mx = src[i];
      if (dst[i] > mx) mx = dst[i];
    }
    GB_REQUIRE(m > 0, "cannot infer node_count from an empty edge list");
    GB_REQUIRE(mx < 0xFFFFFFFFu, "node id 2^32-1 leaves no room for node_count");
    n = mx + 1;
  }
  gb_graph* g = nullptr;
  GB_TRY(new_graph(device, kind, n, &g));
  gb_status st = [&]() -> gb_status {
    DevBuf<uint32_t> d_src, d_dst;
    DevBuf<float> d_w;
    DevBuf<unsigned int> bad;
    GB_TRY(d_src.alloc(m));
    GB_TRY(d_dst.alloc(m));
    GB_TRY(bad.alloc(1));
    GB_CUDA(cudaMemsetAsync(bad.p, 0, 4, g->stream));
    if (m) {
      GB_CUDA(cudaMemcpyAsync(d_src.p, src, m * 4, cudaMemcpyHostToDevice, g->stream));
      GB_CUDA(cudaMemcpyAsync(d_dst.p, dst, m * 4, cudaMemcpyHostToDevice, g->stream));
      k_check_ids<<<grid_for(m, 256), 256, 0, g->stream>>>(d_src.p, m, n, bad.p);
      k_check_ids<<<grid_for(m, 256), 256, 0, g->stream>>>(d_dst.p, m, n, bad.p);
    }
    if (w && kind == GB_KIND_DIRECTED) {
      GB_TRY(d_w.alloc(m));
      if (m) GB_CUDA(cudaMemcpyAsync(d_w.p, w, m * 4, cudaMemcpyHostToDevice, g->stream));
    }
    unsigned int nbad = 0;
    GB_CUDA(cudaMemcpyAsync(&nbad, bad.p, 4, cudaMemcpyDeviceToHost, g->stream));
    GB_CUDA(cudaStreamSynchronize(g->stream));
    GB_REQUIRE(nbad == 0, "%u edge endpoints are >= node_count %u", nbad, n);
    return graph_from_device_edges(g, d_src.p, d_dst.p, d_w.p, m, layout);
  }();
  if (st != GB_OK) {
    gb_graph_free(g);
    return st;
  }
  *out = g;
  return GB_OK;
}

static gb_status rmat_graph(int device, gb_graph_kind kind, uint32_t scale, uint32_t edge_factor,
                            uint64_t seed, gb_layout layout, int weights, gb_graph** out) {
  GB_REQUIRE(out != nullptr, "graph out-pointer is NULL");
  GB_REQUIRE(scale >= 1 && scale <= 31, "scale %u out of range [1,31]", scale);
  GB_REQUIRE(edge_factor >= 1, "edge_factor must be >= 1");
  uint64_t m = (uint64_t)edge_factor << scale;
  uint64_t cap = (kind == GB_KIND_UNDIRECTED) ? 2 * m : m;
  GB_REQUIRE(cap < 0xFFFFFFFFull, "2^%u * %u edges do not fit u32 offsets", scale, edge_factor);
  gb_graph* g = nullptr;
  GB_TRY(new_graph(device, kind, 1u << scale, &g));
  gb_status st = [&]() -> gb_status {
    DevBuf<uint32_t> d_src, d_dst;
    DevBuf<float> d_w;
    GB_TRY(d_src.alloc(m));
    GB_TRY(d_dst.alloc(m));
    k_rmat<<<grid_for(m, 256), 256, 0, g->stream>>>(scale, seed, 0, m, d_src.p, d_dst.p);
    if (weights && kind == GB_KIND_DIRECTED) {
      GB_TRY(d_w.alloc(m));
      k_rmat_weights<<<grid_for(m, 256), 256, 0, g->stream>>>(seed, m, d_w.p);
    }
    GB_CUDA(cudaGetLastError());
    return graph_from_device_edges(g, d_src.p, d_dst.p, d_w.p, m, layout);
  }();
  if (st != GB_OK) {
    gb_graph_free(g);
    return st;
  }
  *out = g;
  return GB_OK;
}

}  // namespace gb

using namespace gb;

// ---- C ABI ---------------------------------------------------------------------------------
extern "C" {

int gb_abi_version(void) { return GB_ABI_VERSION; }
const char* gb_last_error(void) { return last_error().c_str(); }
int gb_device_count(void) {
  int c = 0;
  if (cudaGetDeviceCount(&c) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return c;
}
void gb_set_profiling(int on) { gb::g_profiling = on != 0; }

gb_status gb_digraph_from_csr_u32(int device, uint32_t n, const uint32_t* out_off,
                                  const uint32_t* out_tgt, const float* out_w, const uint32_t* in_off,
                                  const uint32_t* in_tgt, gb_graph** graph) {
  GB_REQUIRE(graph != nullptr, "graph out-pointer is NULL");
  GB_REQUIRE(n > 0, "node_count must be > 0");
  GB_TRY(validate_host_csr(n, out_off, out_tgt, "out"));
  GB_TRY(validate_host_csr(n, in_off, in_tgt, "in"));
  GB_REQUIRE(out_off[n] == in_off[n], "out and in CSR disagree on the edge count");
  gb_graph* g = nullptr;
  GB_TRY(new_graph(device, GB_KIND_DIRECTED, n, &g));
  gb_status st = upload_csr(g->stream, n, out_off, out_tgt, out_w, &g->out);
  if (st == GB_OK) st = upload_csr(g->stream, n, in_off, in_tgt, nullptr, &g->in);
  if (st == GB_OK) st = validate_device_targets(g->stream, n, g->out, "out");
  if (st == GB_OK) st = validate_device_targets(g->stream, n, g->in, "in");
  if (st == GB_OK && cudaStreamSynchronize(g->stream) != cudaSuccess)
    st = fail(GB_ERR_CUDA, "upload failed: %s", cudaGetErrorString(cudaGetLastError()));
  if (st != GB_OK) {
    gb_graph_free(g);
    return st;
  }
  *graph = g;
  return GB_OK;
}

gb_status gb_graph_from_csr_u32(int device, uint32_t n, const uint32_t* off, const uint32_t* tgt,
                                gb_graph** graph) {
  GB_REQUIRE(graph != nullptr, "graph out-pointer is NULL");
  GB_REQUIRE(n > 0, "node_count must be > 0");
  GB_TRY(validate_host_csr(n, off, tgt, "undirected"));
  gb_graph* g = nullptr;
  GB_TRY(new_graph(device, GB_KIND_UNDIRECTED, n, &g));
  gb_status st = upload_csr(g->stream, n, off, tgt, nullptr, &g->out);
  if (st == GB_OK) st = validate_device_targets(g->stream, n, g->out, "undirected");
  if (st == GB_OK && cudaStreamSynchronize(g->stream) != cudaSuccess)
    st = fail(GB_ERR_CUDA, "upload failed: %s", cudaGetErrorString(cudaGetLastError()));
  if (st != GB_OK) {
    gb_graph_free(g);
    return st;
  }
  *graph = g;
  return GB_OK;
}

gb_status gb_digraph_from_edges_u32(int device, const uint32_t* src, const uint32_t* dst,
                                    const float* weights, uint64_t m, uint32_t n, gb_layout layout,
                                    gb_graph** graph) {
  return graph_from_host_edges(device, GB_KIND_DIRECTED, src, dst, weights, m, n, layout, graph);
}

gb_status gb_graph_from_edges_u32(int device, const uint32_t* src, const uint32_t* dst, uint64_t m,
                                  uint32_t n, gb_layout layout, gb_graph** graph) {
  return graph_from_host_edges(device, GB_KIND_UNDIRECTED, src, dst, nullptr, m, n, layout, graph);
}

gb_status gb_digraph_rmat(int device, uint32_t scale, uint32_t edge_factor, uint64_t seed,
                          gb_layout layout, int weights, gb_graph** graph) {
  return rmat_graph(device, GB_KIND_DIRECTED, scale, edge_factor, seed, layout, weights, graph);
}
gb_status gb_graph_rmat(int device, uint32_t scale, uint32_t edge_factor, uint64_t seed,
                        gb_layout layout, gb_graph** graph) {
  return rmat_graph(device, GB_KIND_UNDIRECTED, scale, edge_factor, seed, layout, 0, graph);
}

gb_status gb_rmat_edges(int device, uint32_t scale, uint64_t seed, uint64_t first, uint64_t count,
                        uint32_t* src, uint32_t* dst) {
  GB_REQUIRE(scale >= 1 && scale <= 31, "scale %u out of range [1,31]", scale);
  GB_REQUIRE(count == 0 || (src && dst), "output arrays are NULL");
  int devs = gb_device_count();
  if (devs <= 0) return fail(GB_ERR_CUDA, "no CUDA device available: libgraph_b200 has no CPU fallback");
  GB_REQUIRE(device >= 0 && device < devs, "device %d out of range", device);
  DeviceGuard guard(device);
  DevBuf<uint32_t> d_src, d_dst;
  GB_TRY(d_src.alloc(count));
  GB_TRY(d_dst.alloc(count));
  if (count) {
    k_rmat<<<grid_for(count, 256), 256>>>(scale, seed, first, count, d_src.p, d_dst.p);
    GB_CUDA(cudaGetLastError());
    GB_CUDA(cudaMemcpy(src, d_src.p, count * 4, cudaMemcpyDeviceToHost));
    GB_CUDA(cudaMemcpy(dst, d_dst.p, count * 4, cudaMemcpyDeviceToHost));
  }
  return GB_OK;
}

gb_status gb_graph_free(gb_graph* g) {
  if (!g) return GB_OK;
  DeviceGuard guard(g->device);
  if (g->stream) cudaStreamSynchronize(g->stream);
  if (g->pr_plan) free_pr_plan(g->pr_plan);
  g->out = DevCsr();
  g->in = DevCsr();
  if (g->ev_begin) cudaEventDestroy(g->ev_begin);
  if (g->ev_end) cudaEventDestroy(g->ev_end);
  if (g->stream) cudaStreamDestroy(g->stream);
  delete g;
  return GB_OK;
}

gb_status gb_graph_get_info(const gb_graph* g, gb_graph_info* info) {
  GB_REQUIRE(g && info, "NULL argument");
  info->kind = (uint32_t)g->kind;
  info->node_count = g->n;
  info->target_count = g->out.len;
  info->edge_count = (g->kind == GB_KIND_UNDIRECTED) ? g->out.len / 2 : g->out.len;
  info->has_weights = g->out.w.p != nullptr;
  info->device = g->device;
  info->device_bytes = g->out.bytes() + g->in.bytes() + pr_plan_bytes(g->pr_plan);
  return GB_OK;
}

static const DevCsr* pick_csr(const gb_graph* g, gb_csr_which which) {
  if (g->kind == GB_KIND_DIRECTED) {
    if (which == GB_CSR_OUT) return &g->out;
    if (which == GB_CSR_IN) return &g->in;
    return nullptr;
  }
  return which == GB_CSR_UNDIRECTED ? &g->out : nullptr;
}

gb_status gb_graph_csr_len(const gb_graph* g, gb_csr_which which, uint64_t* len) {
  GB_REQUIRE(g && len, "NULL argument");
  const DevCsr* c = pick_csr(g, which);
  if (!c) return fail(GB_ERR_UNSUPPORTED, "graph kind %d has no CSR %d", (int)g->kind, (int)which);
  *len = c->len;
  return GB_OK;
}

gb_status gb_graph_copy_csr(const gb_graph* g, gb_csr_which which, uint32_t* off, uint32_t* tgt,
                            float* w) {
  GB_REQUIRE(g && off, "NULL argument");
  const DevCsr* c = pick_csr(g, which);
  if (!c) return fail(GB_ERR_UNSUPPORTED, "graph kind %d has no CSR %d", (int)g->kind, (int)which);
  DeviceGuard guard(g->device);
  std::lock_guard<std::mutex> lock(g->mu);
  GB_CUDA(cudaMemcpyAsync(off, c->off.p, ((size_t)g->n + 1) * 4, cudaMemcpyDeviceToHost, g->stream));
  if (tgt && c->len) {
    GB_REQUIRE(c->tgt.p != nullptr, "this handle holds no targets for that CSR (page-rank-only twin)");
    GB_CUDA(cudaMemcpyAsync(tgt, c->tgt.p, c->len * 4, cudaMemcpyDeviceToHost, g->stream));
  }
  if (w) {
    GB_REQUIRE(c->w.p != nullptr, "this CSR carries no edge values");
    if (c->len) GB_CUDA(cudaMemcpyAsync(w, c->w.p, c->len * 4, cudaMemcpyDeviceToHost, g->stream));
  }
  GB_CUDA(cudaStreamSynchronize(g->stream));
  return GB_OK;
}

void* gb_graph_stream(const gb_graph* g) { return g ? (void*)g->stream : nullptr; }

gb_status gb_graph_last_timing(const gb_graph* g, gb_timing* t) {
  GB_REQUIRE(g && t, "NULL argument");
  *t = g->timing;
  return GB_OK;
}

// to_undirected: csr.rs:391-464 — every out-edge (u, v) of the directed graph becomes an edge of a
// fresh UndirectedCsrGraph, fed in out-CSR order.
gb_status gb_to_undirected(const gb_graph* dg, gb_layout layout, gb_graph** graph) {
  GB_REQUIRE(dg && graph, "NULL argument");
  if (dg->kind != GB_KIND_DIRECTED) return fail(GB_ERR_UNSUPPORTED, "to_undirected needs a directed graph");
  GB_REQUIRE((int)layout >= 0 && (int)layout <= 2, "bad layout %d", (int)layout);
  uint64_t m = dg->out.len;
  GB_REQUIRE(m == 0 || dg->out.tgt.p != nullptr, "this handle holds no out targets (page-rank-only twin)");
  GB_REQUIRE(2 * m < 0xFFFFFFFFull, "undirected twin would exceed u32 offsets");
  DeviceGuard guard(dg->device);
  gb_graph* g = nullptr;
  GB_TRY(new_graph(dg->device, GB_KIND_UNDIRECTED, dg->n, &g));
  gb_status st = [&]() -> gb_status {
    std::lock_guard<std::mutex> lock(dg->mu);
    DevBuf<uint32_t> rows;
    GB_TRY(rows.alloc(m));
    if (m) k_expand_rows<<<grid_for((uint64_t)dg->n * 32, 256), 256, 0, g->stream>>>(dg->out.off.p, dg->n, m, rows.p);
    GB_CUDA(cudaGetLastError());
    return graph_from_device_edges(g, rows.p, dg->out.tgt.p, nullptr, m, layout);
  }();
  if (st != GB_OK) {
    gb_graph_free(g);
    return st;
  }
  *graph = g;
  return GB_OK;
}

// make_degree_ordered: graph_ops.rs:511-638.  (degree, id) pairs sorted descending (:555), so
// ties give the larger old id the smaller new id; rows rewritten and sorted ascending (:629).
gb_status gb_make_degree_ordered(gb_graph* g) {
  GB_REQUIRE(g != nullptr, "NULL graph");
  if (g->kind != GB_KIND_UNDIRECTED)
    return fail(GB_ERR_UNSUPPORTED, "make_degree_ordered is defined for undirected graphs (graph_ops.rs:240-253)");
  DeviceGuard guard(g->device);
  std::lock_guard<std::mutex> lock(g->mu);
  cudaStream_t s = g->stream;
  const uint32_t n = g->n;
  const uint64_t len = g->out.len;
  const uint32_t bits = bits_for(n);
  DevBuf<uint64_t> dk, dk_alt;
  GB_TRY(dk.alloc(n));
  GB_TRY(dk_alt.alloc(n));
  k_degree_keys<<<grid_for(n, 256), 256, 0, s>>>(g->out.off.p, n, dk.p);
  cub::DoubleBuffer<uint64_t> db(dk.p, dk_alt.p);
  size_t tmp_bytes = 0;
  GB_CUDA(cub::DeviceRadixSort::SortKeysDescending(nullptr, tmp_bytes, db, (int)n, 0, 64, s));
  DevBuf<uint8_t> tmp;
  GB_TRY(tmp.alloc(tmp_bytes));
  GB_CUDA(cub::DeviceRadixSort::SortKeysDescending(tmp.p, tmp_bytes, db, (int)n, 0, 64, s));
  DevBuf<uint32_t> new_id, new_deg, rows;
  GB_TRY(new_id.alloc(n));
  GB_TRY(new_deg.alloc(n));
  k_rank_to_newid<<<grid_for(n, 256), 256, 0, s>>>(db.Current(), n, new_id.p, new_deg.p);
  DevCsr fresh;
  GB_TRY(fresh.off.alloc((size_t)n + 1));
  GB_TRY(fresh.tgt.alloc(len, 8));
  fresh.len = len;
  GB_CUDA(cudaMemsetAsync(fresh.tgt.p + len, 0, 8 * 4, s));
  if (len) {
    GB_TRY(rows.alloc(len));
    k_expand_rows<<<grid_for((uint64_t)n * 32, 256), 256, 0, s>>>(g->out.off.p, n, len, rows.p);
    DevBuf<uint64_t> keys, keys_alt;
    GB_TRY(keys.alloc(len));
    GB_TRY(keys_alt.alloc(len));
    k_relabel_keys<<<grid_for(len, 256), 256, 0, s>>>(rows.p, g->out.tgt.p, new_id.p, len, bits, keys.p);
    cub::DoubleBuffer<uint64_t> kb(keys.p, keys_alt.p);
    size_t sb = 0;
    GB_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, sb, kb, len, 0, (int)(2 * bits), s));
    DevBuf<uint8_t> stmp;
    GB_TRY(stmp.alloc(sb));
    GB_CUDA(cub::DeviceRadixSort::SortKeys(stmp.p, sb, kb, len, 0, (int)(2 * bits), s));
    k_unpack_targets<<<grid_for(len, 256), 256, 0, s>>>(kb.Current(), len, bits, fresh.tgt.p);
    GB_TRY(offsets_from_sorted(s, RowOfKey{kb.Current(), bits}, len, n, fresh.off.p));
    GB_CUDA(cudaGetLastError());
    GB_CUDA(cudaStreamSynchronize(s));
  } else {
    GB_CUDA(cudaMemsetAsync(fresh.off.p, 0, ((size_t)n + 1) * 4, s));
    GB_CUDA(cudaStreamSynchronize(s));
  }
  g->out = std::move(fresh);  // SwapCsr::swap_csr, csr.rs:120-122
  return GB_OK;
}

// in_degree_partition: graph_ops.rs:431-439 + greedy_node_map_partition :479-509
gb_status gb_in_degree_partition(const gb_graph* g, uint32_t parts, uint32_t* ranges) {
  GB_REQUIRE(g && ranges, "NULL argument");
  GB_REQUIRE(parts >= 1, "parts must be >= 1");
  if (g->kind != GB_KIND_DIRECTED) return fail(GB_ERR_UNSUPPORTED, "in_degree_partition needs a directed graph");
  DeviceGuard guard(g->device);
  std::vector<uint32_t> off((size_t)g->n + 1);
  {
    std::lock_guard<std::mutex> lock(g->mu);
    GB_CUDA(cudaMemcpyAsync(off.data(), g->in.off.p, off.size() * 4, cudaMemcpyDeviceToHost, g->stream));
    GB_CUDA(cudaStreamSynchronize(g->stream));
  }
  const uint64_t m = g->in.len;
  const uint64_t batch = (m + parts - 1) / parts;  // ceil(m / parts)
  uint32_t count = 0;
  uint64_t acc = 0;
  ranges[0] = 0;
  for (uint32_t v = 0; v < g->n; ++v) {
    acc += off[v + 1] - off[v];
    if ((count < parts - 1 && acc >= batch) || v == g->n - 1) {
      ranges[++count] = v + 1;
      acc = 0;
    }
  }
  for (uint32_t i = count + 1; i <= parts; ++i) ranges[i] = g->n;  // unused trailing ranges are empty
  return GB_OK;
}

}  // extern "C"
