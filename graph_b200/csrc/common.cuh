// common.cuh — shared plumbing of libgraph_b200.so (error handling, device buffers, graph handle).
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "../../include/graph_b200.h"

namespace gb {

// ---- thread-local error message behind gb_last_error() ---------------------------------------
std::string& last_error();
gb_status fail(gb_status st, const char* fmt, ...);

#define GB_CUDA(expr)                                                                          \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      gb_status _s = (_e == cudaErrorMemoryAllocation) ? GB_ERR_OOM : GB_ERR_CUDA;             \
      return gb::fail(_s, "%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
    }                                                                                          \
  } while (0)

#define GB_TRY(expr)                 \
  do {                               \
    gb_status _s = (expr);           \
    if (_s != GB_OK) return _s;      \
  } while (0)

#define GB_REQUIRE(cond, ...)                                   \
  do {                                                          \
    if (!(cond)) return gb::fail(GB_ERR_INVALID, __VA_ARGS__);  \
  } while (0)

// ---- device buffer (RAII) ------------------------------------------------------------------------
// Stream-ordered allocations from the device's default memory pool, whose release threshold is raised
// once so that freed blocks stay cached: the one-shot entry points (gb_page_rank_csr_u32 uploads a CSR,
// builds a layout and frees everything on every call) then pay for their ~9 GB of cudaMalloc/cudaFree
// only the first time.  Semantics stay those of cudaMalloc/cudaFree: alloc returns memory usable on
// any stream at once, release waits for the device before the block goes back to the pool.
inline void devbuf_pool_setup() {
  static thread_local int configured_for = -1;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev == configured_for) return;
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
    unsigned long long keep = ~0ull;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
  }
  configured_for = dev;
}
// Inside a DevBufStreamScope every buffer released by this thread is known to have been used on that
// one stream only, so release waits for the stream instead of the device: the layout build can then drop
// its temporaries while a copy stream is still bringing in the rest of the graph.
inline cudaStream_t*& devbuf_scope_slot() {
  static thread_local cudaStream_t* slot = nullptr;
  return slot;
}
struct DevBufStreamScope {
  cudaStream_t stream;
  cudaStream_t* prev;
  explicit DevBufStreamScope(cudaStream_t s) : stream(s), prev(devbuf_scope_slot()) { devbuf_scope_slot() = &stream; }
  ~DevBufStreamScope() { devbuf_scope_slot() = prev; }
  DevBufStreamScope(const DevBufStreamScope&) = delete;
  DevBufStreamScope& operator=(const DevBufStreamScope&) = delete;
};
template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
  DevBuf& operator=(DevBuf&& o) noexcept {
    if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; }
    return *this;
  }
  ~DevBuf() { release(); }
  void release() {
    if (p) {
      // like cudaFree: nothing may still be using the block
      if (cudaStream_t* scoped = devbuf_scope_slot()) cudaStreamSynchronize(*scoped);
      else cudaDeviceSynchronize();
      cudaFreeAsync(p, cudaStreamLegacy);
    }
    p = nullptr;
    n = 0;
  }
  // allocates count elements (+ pad elements of slack so 128-bit loads may overrun the tail)
  gb_status alloc(size_t count, size_t pad = 0) {
    release();
    size_t bytes = (count + pad) * sizeof(T);
    if (bytes == 0) bytes = sizeof(T);
    devbuf_pool_setup();
    cudaError_t e = cudaMallocAsync(reinterpret_cast<void**>(&p), bytes, cudaStreamLegacy);
    if (e == cudaSuccess) e = cudaStreamSynchronize(cudaStreamLegacy);  // usable from every stream now
    if (e != cudaSuccess) {
      p = nullptr;
      cudaGetLastError();
      return fail(e == cudaErrorMemoryAllocation ? GB_ERR_OOM : GB_ERR_CUDA,
                  "cudaMallocAsync(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
    }
    n = count;
    return GB_OK;
  }
  size_t bytes() const { return n * sizeof(T); }
};

// ---- device CSR ----------------------------------------------------------------------------
// offsets[n+1] u32, targets[len] u32 (+8 entries of zeroed slack for 128-bit loads), optional
// SoA weights[len] f32 (the host API exposes the reference's 8-byte AoS Target{u32,f32}).
struct DevCsr {
  DevBuf<uint32_t> off;
  DevBuf<uint32_t> tgt;
  DevBuf<float> w;
  uint64_t len = 0;
  bool present() const { return off.p != nullptr; }
  uint64_t bytes() const { return off.bytes() + tgt.bytes() + w.bytes(); }
};

struct PrPlan;  // pagerank.cu

}  // namespace gb

// the opaque handle of the C ABI
namespace gb {
// Host targets of the in-CSR that are still on their way to the device (gb_page_rank_csr_u32): the copy
// stream brings them in row-aligned chunks, and the layout build classifies chunk k while chunk k+1 is
// on the bus.  row_begin[k] .. row_begin[k+1] are the ORIGINAL row ids of chunk k.
struct TargetFeed {
  std::vector<uint32_t> row_begin;   // [chunks + 1]
  std::vector<uint64_t> edge_begin;  // [chunks + 1] = in_off[row_begin[k]]
  std::vector<cudaEvent_t> ready;    // [chunks] recorded on the copy stream behind each chunk
};
}  // namespace gb

struct gb_graph {
  int device = 0;
  gb_graph_kind kind = GB_KIND_DIRECTED;
  uint32_t n = 0;
  gb::DevCsr out;  // directed: csr_out; undirected: the single csr
  gb::DevCsr in;   // directed only: csr_inc
  cudaStream_t stream = nullptr;
  cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
  mutable std::mutex mu;            // algorithms on one handle serialise on its stream
  mutable gb::PrPlan* pr_plan = nullptr;  // lazily built PageRank layout (pagerank.cu)
  mutable const gb::TargetFeed* feed = nullptr;  // set only while gb_page_rank_csr_u32 streams the targets in
  mutable gb_timing timing{};
  uint64_t extra_bytes = 0;
};

namespace gb {
void free_pr_plan(PrPlan* p);
uint64_t pr_plan_bytes(const PrPlan* p);
bool profiling_on();

// RAII: make the graph's device current for the duration of a call
struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != dev) ok = (cudaSetDevice(dev) == cudaSuccess);
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

// CSR construction on device (graph.cu)
// rows/cols: device arrays of `count` entries (consumed / overwritten). Builds `csr` with the
// given layout; n rows.  w may be null.
gb_status build_csr_device(cudaStream_t s, uint32_t n, uint32_t* d_rows, uint32_t* d_cols, float* d_w,
                           uint64_t count, gb_layout layout, DevCsr* csr);
gb_status new_graph(int device, gb_graph_kind kind, uint32_t n, gb_graph** out);
// uploads a host CSR (offsets always, targets/weights when non-null) and validates it on the device
gb_status upload_host_csr(cudaStream_t s, uint32_t n, const uint32_t* off, const uint32_t* tgt, const float* w,
                          DevCsr* csr, const char* what);

inline unsigned grid_for(uint64_t items, unsigned block, unsigned max_blocks = 148u * 16u) {
  uint64_t b = (items + block - 1) / block;
  if (b < 1) b = 1;
  if (b > max_blocks) b = max_blocks;
  return static_cast<unsigned>(b);
}
}  // namespace gb
