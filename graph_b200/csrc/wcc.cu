// wcc.cu — weakly connected components (Afforest) on the device CSR pair.
//
// Replaces crates/algos/src/wcc.rs:127-139,158-301 (`wcc_afforest`, `sample_subgraph`,
// `find_largest_component`, `link_remaining`) and crates/algos/src/afforest.rs:22-56,100-114
// (`Afforest::union/compress/to_vec`).  Same five phases, same link rule (hook the higher root
// under the lower with a CAS), so parent[x] <= x holds throughout and after the final compress
// every entry is the minimum node id of its component — the value `to_vec()` returns.
//
// Algorithmic bytes per run: 8m + 16n + 8 (both CSRs once, parent read + write); the sampling
// phase lets most vertices skip their edge lists, so effective GB/s can exceed the HBM peak.
#include <algorithm>
#include <vector>

#include "common.cuh"

namespace gb {

__device__ __forceinline__ uint32_t ld_parent(const uint32_t* p, uint32_t i) {
  return *((const volatile uint32_t*)(p + i));
}

// Afforest::union, afforest.rs:22-39
__device__ __forceinline__ void af_link(uint32_t* parent, uint32_t u, uint32_t v) {
  uint32_t p1 = ld_parent(parent, u);
  uint32_t p2 = ld_parent(parent, v);
  while (p1 != p2) {
    const uint32_t high = p1 > p2 ? p1 : p2;
    const uint32_t low = p1 + p2 - high;
    const uint32_t p_high = ld_parent(parent, high);
    if (p_high == low) break;
    if (p_high == high && atomicCAS(parent + high, high, low) == high) break;
    p1 = ld_parent(parent, ld_parent(parent, high));
    p2 = ld_parent(parent, low);
  }
}

__global__ void k_cc_init(uint32_t* __restrict__ parent, uint32_t n) {
  for (uint32_t v = blockIdx.x * blockDim.x + threadIdx.x; v < n; v += gridDim.x * blockDim.x) parent[v] = v;
}

// sample_subgraph, wcc.rs:186-204: link u with its first `rounds` out-neighbours
__global__ void k_cc_sample(const uint32_t* __restrict__ off, const uint32_t* __restrict__ tgt, uint32_t vb,
                            uint32_t n, uint32_t rounds, uint32_t* parent) {
  for (uint32_t u = vb + blockIdx.x * blockDim.x + threadIdx.x; u < n; u += gridDim.x * blockDim.x) {
    const uint32_t b = off[u], e = off[u + 1];
    const uint32_t lim = (e - b < rounds) ? e : b + rounds;  // out_neighbors(u).take(neighbor_rounds)
    for (uint32_t i = b; i < lim; ++i) af_link(parent, u, tgt[i]);
  }
}

// Afforest::compress, afforest.rs:50-56
__global__ void k_cc_compress(uint32_t* parent, uint32_t n) {
  for (uint32_t x = blockIdx.x * blockDim.x + threadIdx.x; x < n; x += gridDim.x * blockDim.x) {
    uint32_t p = ld_parent(parent, x);
    uint32_t pp = ld_parent(parent, p);
    while (p != pp) {
      parent[x] = pp;
      p = pp;
      pp = ld_parent(parent, p);
    }
  }
}

__global__ void k_cc_sample_labels(const uint32_t* __restrict__ parent, uint32_t n, uint32_t count,
                                   uint64_t seed, uint32_t* __restrict__ out) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
    uint64_t z = seed + 0x9E3779B97F4A7C15ull * (i + 1);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    out[i] = parent[(uint32_t)(z % n)];
  }
}

// link_remaining, wcc.rs:274-301: one warp per vertex outside the sampled giant component
__global__ void k_cc_link_remaining(const uint32_t* __restrict__ out_off, const uint32_t* __restrict__ out_tgt,
                                    const uint32_t* __restrict__ in_off, const uint32_t* __restrict__ in_tgt,
                                    uint32_t vb, uint32_t n, uint32_t rounds, uint32_t skip, int use_skip,
                                    uint32_t* parent) {
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
  // lanes first test 32 consecutive vertices, then the warp serves the survivors one by one
  for (uint32_t base = vb + warp * 32; base < n; base += nwarps * 32) {
    const uint32_t mine = base + lane;
    bool live = mine < n;
    uint32_t ob = 0, oe = 0, ib = 0, ie = 0;
    if (live) {
      // offsets are read lane-parallel (coalesced); vertices without remaining edges (all isolated
      // vertices, ~half of an R-MAT graph) never enter the warp-serial part
      ob = out_off[mine];
      oe = out_off[mine + 1];
      ib = in_off[mine];
      ie = in_off[mine + 1];
      ob = (oe - ob > rounds) ? ob + rounds : oe;
      live = (oe > ob) || (ie > ib);
    }
    if (live && use_skip) live = ld_parent(parent, mine) != skip;
    // short lists are linked by their own lane; long ones are served by the whole warp
    const uint32_t work = (oe - ob) + (ie - ib);
    const bool small = live && work <= 8;
    if (small) {
      for (uint32_t i = ob; i < oe; ++i) af_link(parent, mine, out_tgt[i]);
      for (uint32_t i = ib; i < ie; ++i) af_link(parent, mine, in_tgt[i]);
    }
    unsigned mask = __ballot_sync(0xFFFFFFFFu, live && !small);
    while (mask) {
      const int owner = __ffs(mask) - 1;
      mask &= mask - 1;
      const uint32_t u = base + owner;
      const uint32_t b0 = __shfl_sync(0xFFFFFFFFu, ob, owner), e0 = __shfl_sync(0xFFFFFFFFu, oe, owner);
      const uint32_t b1 = __shfl_sync(0xFFFFFFFFu, ib, owner), e1 = __shfl_sync(0xFFFFFFFFu, ie, owner);
      for (uint32_t i = b0 + lane; i < e0; i += 32) af_link(parent, u, out_tgt[i]);
      for (uint32_t i = b1 + lane; i < e1; i += 32) af_link(parent, u, in_tgt[i]);
    }
  }
}

// the most frequent label among `sampling_size` pseudo-random vertices (fixed seed: every rank of a
// sharded run that holds the same parent array picks the same label)
static gb_status most_frequent_label(cudaStream_t s, const uint32_t* d_parent, uint32_t n, uint64_t sampling_size,
                                     uint32_t* label, int* found) {
  *label = 0;
  *found = 0;
  const uint32_t samples = (uint32_t)std::min<uint64_t>(sampling_size, 1u << 20);
  if (samples == 0 || n == 0) return GB_OK;
  DevBuf<uint32_t> d_samp;
  GB_TRY(d_samp.alloc(samples));
  k_cc_sample_labels<<<grid_for(samples, 256), 256, 0, s>>>(d_parent, n, samples, 0x5DEECE66Dull, d_samp.p);
  std::vector<uint32_t> h(samples);
  GB_CUDA(cudaMemcpyAsync(h.data(), d_samp.p, (size_t)samples * 4, cudaMemcpyDeviceToHost, s));
  GB_CUDA(cudaStreamSynchronize(s));
  std::sort(h.begin(), h.end());
  uint32_t best_cnt = 0, run = 0;
  for (uint32_t i = 0; i < samples; ++i) {
    run = (i > 0 && h[i] == h[i - 1]) ? run + 1 : 1;
    if (run > best_cnt) {
      best_cnt = run;
      *label = h[i];
    }
  }
  *found = 1;
  return GB_OK;
}

// union of two forests over the same vertex set: every tree edge (v, other[v]) of the other forest is
// linked into parent[] with the Afforest rule, so parent[] ends up connecting what either forest connected
__global__ void k_cc_merge(uint32_t* parent, const uint32_t* __restrict__ other, uint32_t n) {
  for (uint32_t v = blockIdx.x * blockDim.x + threadIdx.x; v < n; v += gridDim.x * blockDim.x) {
    const uint32_t o = other[v];
    if (o != v) af_link(parent, v, o);
  }
}

static gb_status wcc_impl(const gb_graph* g, const gb_wcc_config* cfg, uint32_t* d_comp, uint32_t* h_comp) {
  GB_REQUIRE(g && cfg, "NULL argument");
  if (g->kind != GB_KIND_DIRECTED)
    return fail(GB_ERR_UNSUPPORTED, "wcc needs a directed graph (wcc.rs:130: DirectedNeighbors)");
  GB_REQUIRE(g->out.len == 0 || g->out.tgt.p != nullptr, "this handle holds no out targets (page-rank-only twin)");
  DeviceGuard guard(g->device);
  std::lock_guard<std::mutex> lock(g->mu);
  cudaStream_t s = g->stream;
  const uint32_t n = g->n;
  DevBuf<uint32_t> tmp;
  if (!d_comp) {
    GB_TRY(tmp.alloc(n));
    d_comp = tmp.p;
  }
  g->timing = gb_timing{};
  GB_CUDA(cudaEventRecord(g->ev_begin, s));
  const unsigned blk = 256;
  const unsigned grid = grid_for(n, blk);
  const uint32_t rounds = (uint32_t)std::min<uint64_t>(cfg->neighbor_rounds, 0xFFFFFFFFull);
  k_cc_init<<<grid, blk, 0, s>>>(d_comp, n);
  // sample_subgraph, wcc.rs:186-204: every vertex links its first `neighbor_rounds` out-neighbours
  const uint32_t sample_rounds = rounds;
  if (sample_rounds) k_cc_sample<<<grid, blk, 0, s>>>(g->out.off.p, g->out.tgt.p, 0, n, sample_rounds, d_comp);
  k_cc_compress<<<grid, blk, 0, s>>>(d_comp, n);
  g->timing.kernel_launches += 2 + (sample_rounds ? 1 : 0);
  // find_largest_component, wcc.rs:245-271 (which component is skipped never changes the result)
  uint32_t skip = 0;
  int use_skip = 0;
  GB_TRY(most_frequent_label(s, d_comp, n, cfg->sampling_size, &skip, &use_skip));
  if (use_skip) g->timing.kernel_launches += 1;
  k_cc_link_remaining<<<grid_for((uint64_t)n, blk), blk, 0, s>>>(g->out.off.p, g->out.tgt.p, g->in.off.p,
                                                               g->in.tgt.p, 0, n, sample_rounds, skip, use_skip,
                                                               d_comp);
  k_cc_compress<<<grid, blk, 0, s>>>(d_comp, n);
  g->timing.kernel_launches += 2;
  GB_CUDA(cudaGetLastError());
  GB_CUDA(cudaEventRecord(g->ev_end, s));
  if (h_comp) GB_CUDA(cudaMemcpyAsync(h_comp, d_comp, (size_t)n * 4, cudaMemcpyDeviceToHost, s));
  GB_CUDA(cudaStreamSynchronize(s));
  float ms = 0.0f;
  GB_CUDA(cudaEventElapsedTime(&ms, g->ev_begin, g->ev_end));
  g->timing.total_ms = ms;
  return GB_OK;
}

}  // namespace gb

extern "C" {
gb_status gb_wcc(const gb_graph* graph, const gb_wcc_config* config, uint32_t* components) {
  GB_REQUIRE(components != nullptr, "components is NULL");
  return gb::wcc_impl(graph, config, nullptr, components);
}
gb_status gb_wcc_device(const gb_graph* graph, const gb_wcc_config* config, uint32_t* d_components) {
  GB_REQUIRE(d_components != nullptr, "d_components is NULL");
  return gb::wcc_impl(graph, config, d_components, nullptr);
}

// ---- multi-GPU WCC: the phases of wcc() (wcc.rs:158-183) over one rank's vertex range ----------------
gb_status gb_wcc_shard_phase(const gb_graph* g, const gb_wcc_config* cfg, uint32_t phase, uint32_t vertex_begin,
                             uint32_t vertex_end, uint32_t skip_label, int use_skip, uint32_t* d_parent,
                             const uint32_t* d_other, void* cuda_stream) {
  GB_REQUIRE(g && cfg && d_parent, "NULL argument");
  if (g->kind != GB_KIND_DIRECTED)
    return gb::fail(GB_ERR_UNSUPPORTED, "wcc needs a directed graph (wcc.rs:130: DirectedNeighbors)");
  GB_REQUIRE(vertex_begin <= vertex_end && vertex_end <= g->n, "bad vertex range [%u, %u)", vertex_begin, vertex_end);
  GB_REQUIRE(g->out.len == 0 || g->out.tgt.p != nullptr, "this handle holds no out targets (page-rank-only twin)");
  gb::DeviceGuard guard(g->device);
  cudaStream_t s = (cudaStream_t)cuda_stream;
  const uint32_t n = g->n, span = vertex_end - vertex_begin;
  const unsigned blk = 256;
  const uint32_t rounds = (uint32_t)std::min<uint64_t>(cfg->neighbor_rounds, 0xFFFFFFFFull);
  switch (phase) {
    case GB_WCC_INIT:
      gb::k_cc_init<<<gb::grid_for(n, blk), blk, 0, s>>>(d_parent, n);
      break;
    case GB_WCC_SAMPLE:  // sample_subgraph over the rank's vertices
      if (rounds && span)
        gb::k_cc_sample<<<gb::grid_for(span, blk), blk, 0, s>>>(g->out.off.p, g->out.tgt.p, vertex_begin, vertex_end,
                                                               rounds, d_parent);
      break;
    case GB_WCC_COMPRESS:
      gb::k_cc_compress<<<gb::grid_for(n, blk), blk, 0, s>>>(d_parent, n);
      break;
    case GB_WCC_MERGE:  // union with another rank's forest
      GB_REQUIRE(d_other != nullptr, "d_other is NULL");
      gb::k_cc_merge<<<gb::grid_for(n, blk), blk, 0, s>>>(d_parent, d_other, n);
      break;
    case GB_WCC_LINK_REMAINING:  // link_remaining over the rank's vertices, skipping the GLOBAL giant component
      if (span)
        gb::k_cc_link_remaining<<<gb::grid_for((uint64_t)span, blk), blk, 0, s>>>(
            g->out.off.p, g->out.tgt.p, g->in.off.p, g->in.tgt.p, vertex_begin, vertex_end, rounds, skip_label,
            use_skip, d_parent);
      break;
    default:
      return gb::fail(GB_ERR_INVALID, "unknown wcc shard phase %u", phase);
  }
  GB_CUDA(cudaGetLastError());
  return GB_OK;
}

gb_status gb_wcc_sample_label(const gb_graph* g, const gb_wcc_config* cfg, const uint32_t* d_parent,
                              uint32_t* label, int* found, void* cuda_stream) {
  GB_REQUIRE(g && cfg && d_parent && label && found, "NULL argument");
  gb::DeviceGuard guard(g->device);
  return gb::most_frequent_label((cudaStream_t)cuda_stream, d_parent, g->n, cfg->sampling_size, label, found);
}
}
