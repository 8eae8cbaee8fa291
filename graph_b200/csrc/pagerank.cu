// pagerank.cu — PageRank over the device in-CSR.
//
// Replaces crates/algos/src/page_rank.rs:58-168 (`page_rank`, `page_rank_iteration`).
//
// Two schedules (gb_pr_mode, include/graph_b200.h):
//   EXACT  — the reference's sweep as one thread executes it: in place, CSR-order f32 sums, separate
//            multiply/add (no FMA), IEEE division.  One warp walks the vertices in id order; lanes
//            only parallelise the gather loads, the additions stay sequential.  Bit-exact with the
//            reference wherever the reference is deterministic (n <= 16384 = one chunk).
//   JACOBI — the throughput path (double-buffered, deterministic).  Vertices are renumbered internally
//            by in-degree descending, then out-degree descending: hub rows come first, rows of equal
//            length are neighbours, and the most gathered sources sit at the front of out_scores.
//            Rows with more than 256 in-edges are cut into padded 256-edge SEGMENTS, 32 of them
//            interleaved per slice; all other rows live in a SELL-32 layout (32 rows of (almost) equal
//            length per slice).  In both, ONE LANE owns one segment / one row: it streams its targets
//            with coalesced 128-bit loads, gathers 8 out_scores per step and keeps a private sum — no
//            cross-lane reduction anywhere.  Both kernels mirror the first 32 K entries of out_scores
//            in shared memory (more would starve L1 of miss slots: see pr_gather).  A finish kernel
//            adds each hub row's segment partials in order and reduces the sweep error in a fixed
//            order.  Vertices without in-edges are constant after the first sweep and skipped.
//
// Algorithmic bytes per sweep: 4m (targets) + 4(n+1) (offsets) + 5*4n (out_scores read+write,
// scores read+write, out-degree read) = 4m + 24n + 4  (BASELINE.md §3).
#include <cub/cub.cuh>

#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace gb {

constexpr int PR_WARPS = 32;        // warps per CTA of the sweep kernels: one persistent CTA per SM
constexpr int PR_THREADS = PR_WARPS * 32;
constexpr int PR_HOT = 32 * 1024;   // out_scores entries mirrored in shared memory (128 KB; L1 keeps ~96 KB)
constexpr int PR_HOT_MAX = 52 * 1024;  // upper bound of the GB_PR_HOT experiment knob
constexpr int PR_FIN_THREADS = 256;
constexpr uint32_t PR_LONG_DEG = 256;  // rows with more in-edges are cut into segments, the rest go to SELL-32
constexpr uint32_t PR_SEG = 256;       // edges per segment (8 per lane)
constexpr uint32_t PR_MAX_PROFILE_EVENTS = 256;  // sweeps bracketed by CUDA events when profiling is on

// the share of a contiguous range of internal rows (the whole graph on one GPU, or one rank's shard
// of the 1-D edge-cut) in the two layouts
struct PrRange {
  uint32_t row_begin = 0, row_end = 0;    // internal rows [row_begin, row_end), clipped to active rows
  uint32_t long_begin = 0, long_end = 0;  // hub rows of the range
  uint32_t seg_begin = 0, seg_end = 0;    // their segments
  uint32_t slice_begin = 0, slice_end = 0;  // SELL slices of the range
  uint32_t sell_row_end = 0;                // one past the last SELL row of the range
  unsigned grid_seg = 0, grid_sell = 0, grid_fin = 1;
  DevBuf<double> block_err;  // per CTA error partials (SELL CTAs, then finish CTAs)
  DevBuf<double> err_hist;   // error of each sweep of the current batch
  DevBuf<uint32_t> ctrl;     // [0] = done flag (sweep number at which tolerance was met), [1] = ticket
  uint64_t bytes() const { return block_err.bytes() + err_hist.bytes() + ctrl.bytes(); }
};

struct PrPlan {
  uint32_t n = 0;
  uint32_t n_active = 0;  // rows with in-degree > 0 (renumbered to [0, n_active))
  uint32_t n_long = 0;    // rows [0, n_long) have more than PR_LONG_DEG in-edges
  uint64_t m = 0;
  uint32_t num_segs = 0, num_slices = 0;
  DevBuf<uint32_t> new_id;    // old id -> internal id
  DevBuf<uint32_t> off;       // internal in-CSR offsets [n+1] (plan-time and partitioning only)
  DevBuf<uint32_t> outdeg;    // out-degree by internal id [n]
  DevBuf<uint32_t> seg_first; // first segment of each hub row [n_long+1]
  DevBuf<uint4> seg_tgt;      // hub rows' targets, 64 uint4 per segment, tail padded with ~0
  DevBuf<float> partial;      // one partial sum per segment
  DevBuf<uint4> sell;         // SELL-32 targets: slice-major, then 4-edge group, then lane
  DevBuf<uint2> slice_meta;   // per slice: (first uint4 index, uint4 groups per lane)
  DevBuf<float> x[2];         // out_scores ping-pong [n]
  DevBuf<float> scores;       // ranks by internal id [n]
  PrRange all;                // the whole active range (single-GPU path)
  uint32_t hot_count = 0;     // entries of out_scores mirrored in shared memory by the sweep kernels
  size_t smem_bytes = 0;      // dynamic shared memory of the sweep kernels
  std::vector<cudaEvent_t> prof_events;
  uint64_t bytes() const {
    return new_id.bytes() + off.bytes() + outdeg.bytes() + seg_first.bytes() + seg_tgt.bytes() + partial.bytes() +
           sell.bytes() + slice_meta.bytes() + x[0].bytes() + x[1].bytes() + scores.bytes() + all.bytes();
  }
};

void free_pr_plan(PrPlan* p) {
  if (!p) return;
  for (cudaEvent_t e : p->prof_events) cudaEventDestroy(e);
  delete p;
}
uint64_t pr_plan_bytes(const PrPlan* p) { return p ? p->bytes() : 0; }

// ---- small device helpers --------------------------------------------------------------------
__device__ __forceinline__ uint4 ld_stream_u4(const uint32_t* p) {
  uint4 r;  // streamed once per sweep: keep it out of L1 so the gathered vector stays there
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ uint32_t ld_stream_u32(const uint32_t* p) {
  uint32_t r;
  asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
  return r;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
  return v;
}

// ---- plan construction kernels ---------------------------------------------------------------
__global__ void k_perm_keys(const uint32_t* __restrict__ in_off, const uint32_t* __restrict__ out_off,
                            uint32_t n, uint64_t* __restrict__ keys, uint32_t* __restrict__ ids) {
  for (uint32_t v = blockIdx.x * blockDim.x + threadIdx.x; v < n; v += gridDim.x * blockDim.x) {
    uint32_t indeg = in_off[v + 1] - in_off[v];
    uint32_t outdeg = out_off[v + 1] - out_off[v];
    // in-degree descending (rows of similar length become neighbours: SELL slices need no padding and
    // hub rows come first), then out-degree descending (hot sources first inside equal in-degrees).
    // R-MAT's expected in- and out-degree of a vertex coincide, so this is also a hot-first order.
    keys[v] = ((uint64_t)(uint32_t)(~indeg) << 32) | (uint32_t)(~outdeg);
    ids[v] = v;
  }
}
__global__ void k_perm_scatter(const uint32_t* __restrict__ sorted_ids, const uint32_t* __restrict__ out_off,
                               const uint32_t* __restrict__ in_off, uint32_t n, uint32_t* __restrict__ new_id,
                               uint32_t* __restrict__ outdeg, uint32_t* __restrict__ indeg) {
  for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r <= n; r += gridDim.x * blockDim.x) {
    if (r == n) {
      indeg[n] = 0;
      continue;
    }
    uint32_t v = sorted_ids[r];
    new_id[v] = r;
    outdeg[r] = out_off[v + 1] - out_off[v];
    indeg[r] = in_off[v + 1] - in_off[v];
  }
}
// ---- sweep kernels (JACOBI) ------------------------------------------------------------------
struct PrArgs {
  const uint32_t* outdeg;
  const float* x_cur;
  float* x_next;
  float* peer_next[7];  // peer-mapped copies of x_next (fused allgather over NVLink); n_peers used
  uint32_t n_peers;
  uint32_t hot_count;   // entries of x_cur mirrored in shared memory (multiple of 4)
  float* scores;
  // hub rows
  const uint4* seg_tgt;
  const uint32_t* seg_first;
  float* partial;
  uint32_t seg_begin, seg_end;
  uint32_t long_begin, long_end;
  // SELL rows
  const uint4* sell;
  const uint2* slice_meta;
  uint32_t slice_begin, slice_end;
  uint32_t sell_row0;     // row of lane 0 of slice 0 (= n_long)
  uint32_t sell_row_end;  // one past the last SELL row of this range
  // error / stop rule
  double* block_err;
  double* err_hist;
  uint32_t* ctrl;
  uint32_t err_base_fin;  // block_err slots [0, err_base_fin) belong to the SELL CTAs
  float base, damping;
  double tolerance;
  double extra_err;   // closed-form error of the skipped zero-in-degree rows (first sweep only)
  uint32_t sweep;     // index inside the current batch
  uint32_t sweep_no;  // 1-based global sweep number
};

// 8 gathers per lane in straight-line predicated code: ids below hot_n read the shared-memory mirror,
// all others (except the padding id ~0) read global memory through L1 (ld.global.nc).  Measured
// (profiles/r01_sweep_hot_head.txt): per-target if/else made every load wait for a scoreboard slot of
// the previous one; and every pending miss holds an L1 line, so the hot head must leave L1 room —
// at 208 KB of shared memory (16 KB of L1) the sweep ran 2.8x slower than at 128 KB (96 KB of L1);
// ld.global.nc.L1::no_allocate was slower still at every size.
__device__ __forceinline__ void pr_gather(const float* x, uint32_t hot_saddr, uint32_t hot_n,
                                          const uint4& ta, const uint4& tb, float (&v)[8]) {
  const uint32_t t[8] = {ta.x, ta.y, ta.z, ta.w, tb.x, tb.y, tb.z, tb.w};
  float vs[8], vg[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.lt.u32 p, %1, %2;\n\t"
        "mov.f32 %0, 0f00000000;\n\t"
        "@p ld.shared.f32 %0, [%3];\n\t}"
        : "=f"(vs[j])
        : "r"(t[j]), "r"(hot_n), "r"(hot_saddr + 4u * t[j]));
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ge.u32 p, %1, %2;\n\t"
        "setp.ne.and.u32 p, %1, 0xffffffff, p;\n\t"
        "mov.f32 %0, 0f00000000;\n\t"
        "@p ld.global.nc.f32 %0, [%3];\n\t}"
        : "=f"(vg[j])
        : "r"(t[j]), "r"(hot_n), "l"(x + t[j]));
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = vs[j] + vg[j];
}
__device__ __forceinline__ float pr_sum8(const float (&v)[8]) {
  return ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]));
}
__device__ __forceinline__ uint4 pr_ld4(const uint4* p) { return ld_stream_u4(reinterpret_cast<const uint32_t*>(p)); }

// the per-vertex update of page_rank.rs:148-158 with the reference's rounding sequence
template <bool PEERS>
__device__ __forceinline__ double pr_update(uint32_t r, float sum, float old, uint32_t deg, const PrArgs& a) {
  const float nw = __fadd_rn(a.base, __fmul_rn(a.damping, sum));
  a.scores[r] = nw;
  const float xo = __fdiv_rn(nw, (float)deg);
  a.x_next[r] = xo;
  // fused allgather: the finished out_score also goes straight into every peer's next vector
  if (PEERS)
    for (uint32_t p = 0; p < a.n_peers; ++p) a.peer_next[p][r] = xo;
  return fabs((double)__fsub_rn(nw, old));
}

__device__ __forceinline__ void pr_load_hot(float* hot, const float* __restrict__ x, uint32_t hot_n) {
  for (uint32_t i = threadIdx.x * 4; i < hot_n; i += PR_THREADS * 4)
    *reinterpret_cast<float4*>(hot + i) = __ldg(reinterpret_cast<const float4*>(x + i));
  __syncthreads();
}

// ---- hub rows: 256-edge segments, 32 segments per slice, one lane per segment ----------------------
// Divergent 4-byte gathers sustain ~0.95 per clock per SM on B200 whatever the load path
// (profiles/r01_gather_ceiling_microbench.txt), so everything around the gathers is kept minimal.  A
// slice interleaves 32 segments group-major / lane-minor exactly like a SELL slice of width 64: every
// lane streams its own segment with 128-bit coalesced loads and keeps a private sum — there is no
// cross-lane reduction (the shuffle tree of a warp-per-segment version cost a third of its time).
__global__ void __launch_bounds__(PR_THREADS, 1) k_pr_seg(const PrArgs a) {
  extern __shared__ __align__(16) float smem[];
  float* hot = smem;
  if (a.ctrl[0] != 0) return;  // tolerance already met by an earlier sweep of this batch
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* x = a.x_cur;
  const uint32_t hot_n = a.hot_count;
  pr_load_hot(hot, x, hot_n);
  const uint32_t hot_saddr = (uint32_t)__cvta_generic_to_shared(hot);
  const uint32_t stride = gridDim.x * PR_WARPS;
  const uint32_t slice_end = (a.seg_end + 31) / 32;
  constexpr uint32_t W4 = PR_SEG / 4;  // uint4 groups per segment
  for (uint32_t sl = a.seg_begin / 32 + blockIdx.x * PR_WARPS + warp; sl < slice_end; sl += stride) {
    const uint4* base = a.seg_tgt + (size_t)sl * (W4 * 32) + lane;
    uint4 ta = pr_ld4(base), tb = pr_ld4(base + 32);
    float acc = 0.0f;
#pragma unroll 2
    for (uint32_t q = 0; q < W4; q += 2) {
      const uint4 na = (q + 2 < W4) ? pr_ld4(base + (q + 2) * 32) : ta;
      const uint4 nb = (q + 3 < W4) ? pr_ld4(base + (q + 3) * 32) : tb;
      float v[8];
      pr_gather(x, hot_saddr, hot_n, ta, tb, v);
      acc += pr_sum8(v);
      ta = na;
      tb = nb;
    }
    const uint32_t seg = sl * 32 + lane;
    if (seg >= a.seg_begin && seg < a.seg_end) a.partial[seg] = acc;
  }
}

// ---- SELL-32 sweep for rows with at most PR_LONG_DEG in-edges -------------------------------------
// Rows are sorted by in-degree, so the 32 rows of a slice have (almost) the same length: one lane per
// row, no reduction, no per-row offsets.  A lane reads its row four targets at a time (128-bit,
// coalesced: the slice is stored group-major, lane-minor), gathers, and adds in row order.  The next
// slice's first targets and row metadata are requested while the current slice is processed.
template <bool PEERS>
__global__ void __launch_bounds__(PR_THREADS, 1) k_pr_sell(const PrArgs a) {
  extern __shared__ __align__(16) float smem[];
  float* hot = smem;
  __shared__ double warp_err[PR_WARPS];
  if (a.ctrl[0] != 0) return;
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* __restrict__ x = a.x_cur;
  const uint32_t hot_n = a.hot_count;
  pr_load_hot(hot, x, hot_n);
  const uint32_t hot_saddr = (uint32_t)__cvta_generic_to_shared(hot);
  double err = 0.0;
  const uint32_t stride = gridDim.x * PR_WARPS;
  const uint4 pad = make_uint4(~0u, ~0u, ~0u, ~0u);
  uint32_t sidx = a.slice_begin + blockIdx.x * PR_WARPS + warp;
  // pipeline state: metadata of this and the next slice, first two target groups + row data of this one
  uint2 meta = make_uint2(0, 0), nmeta = meta;
  uint4 ta = pad, tb = pad;
  float old = 0.0f;
  uint32_t deg = 1;
  if (sidx < a.slice_end) {
    meta = __ldg(a.slice_meta + sidx);
    if (sidx + stride < a.slice_end) nmeta = __ldg(a.slice_meta + sidx + stride);
    const uint4* base = a.sell + meta.x + lane;
    if (0 < meta.y) ta = pr_ld4(base);
    if (1 < meta.y) tb = pr_ld4(base + 32);
    const uint32_t row = a.sell_row0 + 32 * sidx + lane;
    if (row < a.sell_row_end) {
      old = a.scores[row];
      deg = a.outdeg[row];
    }
  }
  while (sidx < a.slice_end) {
    const uint32_t w4 = meta.y;
    const uint4* base = a.sell + meta.x + lane;
    const uint32_t row = a.sell_row0 + 32 * sidx + lane;
    // next slice: first groups, row data; metadata of the slice after it
    const uint32_t nidx = sidx + stride;
    uint4 nta = pad, ntb = pad;
    float nold = 0.0f;
    uint32_t ndeg = 1;
    uint2 nnmeta = make_uint2(0, 0);
    if (nidx < a.slice_end) {
      const uint4* nbase = a.sell + nmeta.x + lane;
      if (0 < nmeta.y) nta = pr_ld4(nbase);
      if (1 < nmeta.y) ntb = pr_ld4(nbase + 32);
      const uint32_t nrow = a.sell_row0 + 32 * nidx + lane;
      if (nrow < a.sell_row_end) {
        nold = a.scores[nrow];
        ndeg = a.outdeg[nrow];
      }
      if (nidx + stride < a.slice_end) nnmeta = __ldg(a.slice_meta + nidx + stride);
    }
    float acc = 0.0f;
    for (uint32_t q = 0; q < w4; q += 2) {
      // targets of the next two groups are requested before this group's gathers are consumed
      const uint4 na = (q + 2 < w4) ? pr_ld4(base + (q + 2) * 32) : pad;
      const uint4 nb = (q + 3 < w4) ? pr_ld4(base + (q + 3) * 32) : pad;
      float v[8];
      pr_gather(x, hot_saddr, hot_n, ta, tb, v);
      acc += pr_sum8(v);
      ta = na;
      tb = nb;
    }
    if (row < a.sell_row_end) err += pr_update<PEERS>(row, acc, old, deg, a);
    sidx = nidx;
    meta = nmeta;
    nmeta = nnmeta;
    ta = nta;
    tb = ntb;
    old = nold;
    deg = ndeg;
  }
  err = warp_sum(err);
  if (lane == 0) warp_err[warp] = err;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tt = 0.0;
#pragma unroll
    for (int w = 0; w < PR_WARPS; ++w) tt += warp_err[w];
    a.block_err[blockIdx.x] = tt;
  }
}

// ---- finish: hub rows = sum of their segment partials (fixed order), then the sweep error ---------
// One warp per hub row; the last CTA to finish reduces all CTA error partials in a fixed order and
// evaluates the stop rule of page_rank.rs:107 on the device.
template <bool PEERS>
__global__ void __launch_bounds__(PR_FIN_THREADS) k_pr_finish(const PrArgs a) {
  constexpr int FIN_WARPS = PR_FIN_THREADS / 32;
  __shared__ double warp_err[FIN_WARPS];
  __shared__ bool is_last;
  if (a.ctrl[0] != 0) return;
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double err = 0.0;
  const uint32_t nwarps = gridDim.x * FIN_WARPS;
  for (uint32_t r = a.long_begin + blockIdx.x * FIN_WARPS + warp; r < a.long_end; r += nwarps) {
    const uint32_t sb = a.seg_first[r], se = a.seg_first[r + 1];
    float old = 0.0f;
    uint32_t deg = 1;
    if (lane == 0) {
      old = a.scores[r];
      deg = a.outdeg[r];
    }
    float p = 0.0f;
    for (uint32_t j = sb + lane; j < se; j += 32) p += a.partial[j];
    p = warp_sum(p);
    if (lane == 0) err += pr_update<PEERS>(r, p, old, deg, a);
  }
  err = warp_sum(err);
  if (lane == 0) warp_err[warp] = err;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < FIN_WARPS; ++w) t += warp_err[w];
    a.block_err[a.err_base_fin + blockIdx.x] = t;
    __threadfence();
    unsigned ticket = atomicAdd(&a.ctrl[1], 1u);
    is_last = (ticket == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  // fixed-order reduction of all CTA partials (deterministic error)
  const uint32_t total = a.err_base_fin + gridDim.x;
  double t = 0.0;
  for (uint32_t i = threadIdx.x; i < total; i += PR_FIN_THREADS) t += ((volatile double*)a.block_err)[i];
  t = warp_sum(t);
  if (lane == 0) warp_err[warp] = t;
  __syncthreads();
  if (threadIdx.x == 0) {
    double e = a.extra_err;
#pragma unroll
    for (int w = 0; w < FIN_WARPS; ++w) e += warp_err[w];
    a.err_hist[a.sweep] = e;
    a.ctrl[1] = 0;
    if (e < a.tolerance) a.ctrl[0] = a.sweep_no;
  }
}

// ---- plan-time helpers of the two layouts ---------------------------------------------------------
__global__ void k_count_rows(const uint32_t* __restrict__ off, uint32_t n, uint32_t* __restrict__ counts) {
  uint32_t act = 0, lng = 0;
  for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
    const uint32_t d = off[r + 1] - off[r];
    act += d > 0;
    lng += d > PR_LONG_DEG;
  }
  for (int o = 16; o > 0; o >>= 1) {
    act += __shfl_xor_sync(0xFFFFFFFFu, act, o);
    lng += __shfl_xor_sync(0xFFFFFFFFu, lng, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (act) atomicAdd(counts + 0, act);
    if (lng) atomicAdd(counts + 1, lng);
  }
}
__global__ void k_seg_counts(const uint32_t* __restrict__ off, uint32_t n_long, uint32_t* __restrict__ cnt) {
  for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r <= n_long; r += gridDim.x * blockDim.x)
    cnt[r] = (r < n_long) ? (off[r + 1] - off[r] + PR_SEG - 1) / PR_SEG : 0;
}
// one warp per hub row: copy its (renumbered) sources into its padded segments (slice-interleaved
// layout: segment s lives in slice s/32 as lane s%32; element j of it is component j%4 of group j/4)
__global__ void k_seg_fill(const uint32_t* __restrict__ in_off, const uint32_t* __restrict__ in_tgt,
                           const uint32_t* __restrict__ old_of, const uint32_t* __restrict__ new_id,
                           const uint32_t* __restrict__ seg_first, uint32_t n_long, uint32_t* __restrict__ out) {
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t r = warp; r < n_long; r += nwarps) {
    const uint32_t old = old_of[r];
    const uint32_t b = in_off[old], d = in_off[old + 1] - b;
    const uint32_t s0 = seg_first[r];
    for (uint32_t j = lane; j < d; j += 32) {
      const uint32_t sg = s0 + j / PR_SEG, e = j % PR_SEG;
      const uint64_t idx = (((uint64_t)(sg / 32) * (PR_SEG / 4) + e / 4) * 32 + (sg % 32)) * 4 + (e % 4);
      out[idx] = new_id[in_tgt[b + j]];
    }
  }
}
__global__ void k_sell_widths(const uint32_t* __restrict__ off, uint32_t row0, uint32_t num_slices,
                              uint32_t* __restrict__ units) {
  for (uint32_t sidx = blockIdx.x * blockDim.x + threadIdx.x; sidx < num_slices; sidx += gridDim.x * blockDim.x) {
    const uint32_t r = row0 + 32 * sidx;  // rows are sorted by degree: the first row of a slice is its longest
    units[sidx] = ((off[r + 1] - off[r] + 3) / 4) * 32;  // uint4 entries of the slice
  }
}
__global__ void k_sell_meta(const uint32_t* __restrict__ units, const uint32_t* __restrict__ bases,
                            uint32_t num_slices, uint2* __restrict__ meta) {
  for (uint32_t sidx = blockIdx.x * blockDim.x + threadIdx.x; sidx < num_slices; sidx += gridDim.x * blockDim.x)
    meta[sidx] = make_uint2(bases[sidx], units[sidx] / 32);
}
__global__ void k_sell_fill(const uint32_t* __restrict__ in_off, const uint32_t* __restrict__ in_tgt,
                            const uint32_t* __restrict__ old_of, const uint32_t* __restrict__ new_id,
                            uint32_t row0, uint32_t row_end, uint32_t num_slices, const uint2* __restrict__ meta,
                            uint4* __restrict__ sell) {
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t sidx = warp; sidx < num_slices; sidx += nwarps) {
    const uint2 m = meta[sidx];
    const uint32_t row = row0 + 32 * sidx + lane;
    uint32_t b = 0, d = 0;
    if (row < row_end) {
      const uint32_t old = old_of[row];
      b = in_off[old];
      d = in_off[old + 1] - b;
    }
    for (uint32_t q = 0; q < m.y; ++q) {
      uint4 v = make_uint4(~0u, ~0u, ~0u, ~0u);
      const uint32_t j = 4 * q;
      if (j + 0 < d) v.x = new_id[in_tgt[b + j + 0]];
      if (j + 1 < d) v.y = new_id[in_tgt[b + j + 1]];
      if (j + 2 < d) v.z = new_id[in_tgt[b + j + 2]];
      if (j + 3 < d) v.w = new_id[in_tgt[b + j + 3]];
      sell[m.x + q * 32 + lane] = v;
    }
  }
}

__global__ void k_pr_init(uint32_t n, uint32_t n_active, float init, float base,
                          const uint32_t* __restrict__ outdeg, float* __restrict__ x0,
                          float* __restrict__ x1, float* __restrict__ scores) {
  for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
    float d = (float)outdeg[r];
    x0[r] = __fdiv_rn(init, d);  // page_rank.rs:75-79; +inf for dangling vertices, never gathered
    if (r < n_active) {
      scores[r] = init;
    } else {
      // no in-edges: after the first sweep score == base + damping * 0 == base, for ever
      scores[r] = base;
      x1[r] = __fdiv_rn(base, d);
    }
  }
}
__global__ void k_pr_fill_inactive(uint32_t n, uint32_t n_active, float base,
                                   const uint32_t* __restrict__ outdeg, float* __restrict__ x) {
  for (uint32_t r = n_active + blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x)
    x[r] = __fdiv_rn(base, (float)outdeg[r]);
}
__global__ void k_unpermute(const float* __restrict__ src, const uint32_t* __restrict__ new_id,
                            uint32_t n, float* __restrict__ dst) {
  for (uint32_t v = blockIdx.x * blockDim.x + threadIdx.x; v < n; v += gridDim.x * blockDim.x)
    dst[v] = src[new_id[v]];
}

// ---- EXACT: the reference sweep on one warp ---------------------------------------------------
// page_rank.rs:58-168 with the loop of :142-160 executed in id order.  Lanes fetch 32 gathered values
// at a time; every lane then performs the same sequential f32 additions in CSR order.
__global__ void __launch_bounds__(32) k_pr_exact(const uint32_t* __restrict__ in_off,
                                                 const uint32_t* __restrict__ in_tgt,
                                                 const uint32_t* __restrict__ out_off, uint32_t n,
                                                 uint64_t max_iterations, double tolerance, float damping,
                                                 float* scores, float* out, uint64_t* ran,
                                                 double* error) {
  const uint32_t lane = threadIdx.x;
  const float nf = (float)n;
  const float init = __fdiv_rn(1.0f, nf);
  const float base = __fdiv_rn(__fsub_rn(1.0f, damping), nf);
  for (uint32_t v = lane; v < n; v += 32) {
    out[v] = __fdiv_rn(init, (float)(out_off[v + 1] - out_off[v]));
    scores[v] = init;
  }
  __syncwarp();
  uint64_t it = 0;
  double err = 0.0;
  for (;;) {
    err = 0.0;
    for (uint32_t u = 0; u < n; ++u) {
      const uint32_t b = in_off[u], e = in_off[u + 1];
      float tot = 0.0f;
      for (uint32_t i = b; i < e; i += 32) {
        const uint32_t cnt = min(32u, e - i);
        float val = 0.0f;
        if (lane < cnt) val = ((volatile float*)out)[in_tgt[i + lane]];
        for (uint32_t j = 0; j < cnt; ++j) tot = __fadd_rn(tot, __shfl_sync(0xFFFFFFFFu, val, j));
      }
      const float old = scores[u];
      const float nw = __fadd_rn(base, __fmul_rn(damping, tot));
      err += fabs((double)__fsub_rn(nw, old));
      __syncwarp();
      if (lane == 0) {
        scores[u] = nw;
        ((volatile float*)out)[u] = __fdiv_rn(nw, (float)(out_off[u + 1] - out_off[u]));
      }
      __syncwarp();
    }
    ++it;
    if (err < tolerance || it == max_iterations) break;  // page_rank.rs:107
  }
  if (lane == 0) {
    *ran = it;
    *error = err;
  }
}

// ---- chunking of a row range -------------------------------------------------------------------
static gb_status build_range(const gb_graph* g, const PrPlan* p, uint32_t row_begin, uint32_t row_end,
                             PrRange* r) {
  cudaStream_t s = g->stream;
  if (row_end > p->n_active) row_end = p->n_active;  // rows without in-edges are never swept
  if (row_begin > row_end) row_begin = row_end;
  r->row_begin = row_begin;
  r->row_end = row_end;
  // SELL part: rows [max(row_begin, n_long), row_end) — shard boundaries sit on slice boundaries
  const uint32_t sb = std::max(row_begin, p->n_long), se = std::max(row_end, p->n_long);
  r->slice_begin = r->slice_end = 0;
  r->sell_row_end = se;
  if (sb < se) {  // an empty SELL share (e.g. a range that starts at n_active) needs no alignment
    GB_REQUIRE((sb - p->n_long) % 32 == 0, "shard boundary %u is not on a SELL slice boundary", row_begin);
    r->slice_begin = (sb - p->n_long) / 32;
    r->slice_end = (se - p->n_long + 31) / 32;
  }
  // hub rows of the range and their segments
  r->long_begin = std::min(row_begin, p->n_long);
  r->long_end = std::min(row_end, p->n_long);
  r->seg_begin = r->seg_end = 0;
  if (p->n_long) {
    uint32_t h[2] = {0, 0};
    GB_CUDA(cudaMemcpyAsync(&h[0], p->seg_first.p + r->long_begin, 4, cudaMemcpyDeviceToHost, s));
    GB_CUDA(cudaMemcpyAsync(&h[1], p->seg_first.p + r->long_end, 4, cudaMemcpyDeviceToHost, s));
    GB_CUDA(cudaStreamSynchronize(s));
    r->seg_begin = h[0];
    r->seg_end = h[1];
  }
  int dev_sms = 148;
  GB_CUDA(cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, g->device));
  const uint64_t seg_slices = r->seg_end > r->seg_begin ? (r->seg_end + 31) / 32 - r->seg_begin / 32 : 0;
  const uint64_t want_seg = (seg_slices + PR_WARPS - 1) / PR_WARPS;
  r->grid_seg = (unsigned)std::min<uint64_t>(want_seg, (uint64_t)dev_sms);  // one persistent CTA per SM
  const uint64_t want_sell = ((uint64_t)(r->slice_end - r->slice_begin) + PR_WARPS - 1) / PR_WARPS;
  r->grid_sell = (unsigned)std::min<uint64_t>(want_sell, (uint64_t)dev_sms);
  const uint64_t want_fin = ((uint64_t)(r->long_end - r->long_begin) + (PR_FIN_THREADS / 32) - 1) / (PR_FIN_THREADS / 32);
  r->grid_fin = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(want_fin, (uint64_t)dev_sms * 8));
  const size_t nerr = (size_t)r->grid_sell + r->grid_fin;
  GB_TRY(r->block_err.alloc(nerr));
  GB_CUDA(cudaMemsetAsync(r->block_err.p, 0, nerr * sizeof(double), s));
  GB_TRY(r->err_hist.alloc(64));
  GB_TRY(r->ctrl.alloc(2));
  GB_CUDA(cudaMemsetAsync(r->ctrl.p, 0, 8, s));
  GB_CUDA(cudaStreamSynchronize(s));
  return GB_OK;
}

// ---- plan ------------------------------------------------------------------------------------
static gb_status build_pr_plan(const gb_graph* g, PrPlan** out_plan) {
  cudaStream_t s = g->stream;
  const uint32_t n = g->n;
  const uint64_t m = g->in.len;
  PrPlan* p = new (std::nothrow) PrPlan();
  if (!p) return fail(GB_ERR_OOM, "host allocation failed");
  p->n = n;
  p->m = m;
  gb_status st = [&]() -> gb_status {
    // 1. permutation: in-degree descending, then out-degree descending, then id
    DevBuf<uint32_t> old_of;  // internal id -> original id (plan-time only)
    {
      DevBuf<uint64_t> keys, keys_alt;
      DevBuf<uint32_t> ids, ids_alt;
      GB_TRY(keys.alloc(n));
      GB_TRY(keys_alt.alloc(n));
      GB_TRY(ids.alloc(n));
      GB_TRY(ids_alt.alloc(n));
      k_perm_keys<<<grid_for(n, 256), 256, 0, s>>>(g->in.off.p, g->out.off.p, n, keys.p, ids.p);
      cub::DoubleBuffer<uint64_t> kb(keys.p, keys_alt.p);
      cub::DoubleBuffer<uint32_t> vb(ids.p, ids_alt.p);
      size_t tb = 0;
      GB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, kb, vb, (int)n, 0, 64, s));
      DevBuf<uint8_t> tmp;
      GB_TRY(tmp.alloc(tb));
      GB_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tb, kb, vb, (int)n, 0, 64, s));
      GB_TRY(p->new_id.alloc(n));
      GB_TRY(p->outdeg.alloc(n));
      GB_TRY(p->off.alloc((size_t)n + 1));
      k_perm_scatter<<<grid_for(n, 256), 256, 0, s>>>(vb.Current(), g->out.off.p, g->in.off.p, n, p->new_id.p,
                                                     p->outdeg.p, p->off.p);
      GB_TRY(old_of.alloc(n));
      GB_CUDA(cudaMemcpyAsync(old_of.p, vb.Current(), (size_t)n * 4, cudaMemcpyDeviceToDevice, s));
      GB_CUDA(cudaGetLastError());
      GB_CUDA(cudaStreamSynchronize(s));
    }
    // 2. internal offsets = running sum of the permuted in-degrees.  The rows themselves are never
    //    materialised in internal order: the two layouts below are filled straight from the original
    //    in-CSR through old_of / new_id (a row keeps its original entry order, which fixes the
    //    summation order; no billion-key sort on the build path).
    {
      size_t tb = 0;
      GB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, p->off.p, p->off.p, (int64_t)n + 1, s));
      DevBuf<uint8_t> tmp;
      GB_TRY(tmp.alloc(tb));
      GB_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, p->off.p, p->off.p, (int64_t)n + 1, s));
      GB_CUDA(cudaStreamSynchronize(s));
    }
    // 3. row classes: rows are ordered by in-degree, so both classes are prefixes
    {
      DevBuf<uint32_t> counts;
      GB_TRY(counts.alloc(2));
      GB_CUDA(cudaMemsetAsync(counts.p, 0, 8, s));
      k_count_rows<<<grid_for(n, 256), 256, 0, s>>>(p->off.p, n, counts.p);
      uint32_t h[2] = {0, 0};
      GB_CUDA(cudaMemcpyAsync(h, counts.p, 8, cudaMemcpyDeviceToHost, s));
      GB_CUDA(cudaStreamSynchronize(s));
      p->n_active = h[0];
      p->n_long = h[1];
    }
    // 3b. SELL-32 layout of rows [n_long, n_active)
    p->num_slices = (p->n_active - p->n_long + 31) / 32;
    if (p->num_slices) {
      DevBuf<uint32_t> units, bases;
      GB_TRY(units.alloc(p->num_slices));
      GB_TRY(bases.alloc(p->num_slices));
      k_sell_widths<<<grid_for(p->num_slices, 256), 256, 0, s>>>(p->off.p, p->n_long, p->num_slices, units.p);
      size_t tb = 0;
      GB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, units.p, bases.p, (int)p->num_slices, s));
      DevBuf<uint8_t> tmp;
      GB_TRY(tmp.alloc(tb));
      GB_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, units.p, bases.p, (int)p->num_slices, s));
      uint32_t last_base = 0, last_units = 0;
      GB_CUDA(cudaMemcpyAsync(&last_base, bases.p + p->num_slices - 1, 4, cudaMemcpyDeviceToHost, s));
      GB_CUDA(cudaMemcpyAsync(&last_units, units.p + p->num_slices - 1, 4, cudaMemcpyDeviceToHost, s));
      GB_CUDA(cudaStreamSynchronize(s));
      const size_t total_units = (size_t)last_base + last_units;
      GB_TRY(p->slice_meta.alloc(p->num_slices));
      GB_TRY(p->sell.alloc(total_units, 64));
      k_sell_meta<<<grid_for(p->num_slices, 256), 256, 0, s>>>(units.p, bases.p, p->num_slices, p->slice_meta.p);
      k_sell_fill<<<grid_for((uint64_t)p->num_slices * 32, 256), 256, 0, s>>>(
          g->in.off.p, g->in.tgt.p, old_of.p, p->new_id.p, p->n_long, p->n_active, p->num_slices,
          p->slice_meta.p, p->sell.p);
      GB_CUDA(cudaGetLastError());
      GB_CUDA(cudaStreamSynchronize(s));
    }
    // 3c. hub rows: padded 256-edge segments
    p->num_segs = 0;
    GB_TRY(p->seg_first.alloc((size_t)p->n_long + 1));
    {
      k_seg_counts<<<grid_for((uint64_t)p->n_long + 1, 256), 256, 0, s>>>(p->off.p, p->n_long, p->seg_first.p);
      size_t tb = 0;
      GB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, p->seg_first.p, p->seg_first.p, (int)(p->n_long + 1), s));
      DevBuf<uint8_t> tmp;
      GB_TRY(tmp.alloc(tb));
      GB_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, p->seg_first.p, p->seg_first.p, (int)(p->n_long + 1), s));
      GB_CUDA(cudaMemcpyAsync(&p->num_segs, p->seg_first.p + p->n_long, 4, cudaMemcpyDeviceToHost, s));
      GB_CUDA(cudaStreamSynchronize(s));
    }
    const size_t seg_slices = ((size_t)p->num_segs + 31) / 32;
    GB_TRY(p->seg_tgt.alloc(seg_slices * 32 * (PR_SEG / 4), 64));
    GB_CUDA(cudaMemsetAsync(p->seg_tgt.p, 0xFF, (seg_slices * 32 * (PR_SEG / 4) + 64) * sizeof(uint4), s));  // ~0 = padding
    GB_TRY(p->partial.alloc(std::max<size_t>(p->num_segs, 1)));
    if (p->num_segs) {
      k_seg_fill<<<grid_for((uint64_t)p->n_long * 32, 256), 256, 0, s>>>(
          g->in.off.p, g->in.tgt.p, old_of.p, p->new_id.p, p->seg_first.p, p->n_long,
          reinterpret_cast<uint32_t*>(p->seg_tgt.p));
      GB_CUDA(cudaGetLastError());
    }
    GB_CUDA(cudaStreamSynchronize(s));
    uint32_t hot_cap = PR_HOT;
    if (const char* e = getenv("GB_PR_HOT")) hot_cap = std::min<uint32_t>((uint32_t)PR_HOT_MAX, (uint32_t)atoi(e)) & ~3u;
    p->hot_count = std::min<uint32_t>(hot_cap, n & ~3u);
    p->smem_bytes = (size_t)p->hot_count * sizeof(float) + 16;
    GB_CUDA(cudaFuncSetAttribute(k_pr_seg, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem_bytes));
    GB_CUDA(cudaFuncSetAttribute(k_pr_sell<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem_bytes));
    GB_CUDA(cudaFuncSetAttribute(k_pr_sell<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem_bytes));
    // 4. chunking of the whole active range + state vectors
    GB_TRY(build_range(g, p, 0, p->n_active, &p->all));
    GB_TRY(p->x[0].alloc(n));
    GB_TRY(p->x[1].alloc(n));
    GB_TRY(p->scores.alloc(n));
    GB_CUDA(cudaStreamSynchronize(s));
    return GB_OK;
  }();
  if (st != GB_OK) {
    free_pr_plan(p);
    return st;
  }
  *out_plan = p;
  return GB_OK;
}

static PrArgs make_args(const PrPlan* p, const PrRange* rg, float base, float damping, double tolerance) {
  PrArgs a{};
  a.outdeg = p->outdeg.p;
  a.seg_tgt = p->seg_tgt.p;
  a.seg_first = p->seg_first.p;
  a.partial = p->partial.p;
  a.seg_begin = rg->seg_begin;
  a.seg_end = rg->seg_end;
  a.long_begin = rg->long_begin;
  a.long_end = rg->long_end;
  a.sell = p->sell.p;
  a.slice_meta = p->slice_meta.p;
  a.slice_begin = rg->slice_begin;
  a.slice_end = rg->slice_end;
  a.sell_row0 = p->n_long;
  a.sell_row_end = rg->sell_row_end;
  a.block_err = rg->block_err.p;
  a.err_hist = rg->err_hist.p;
  a.ctrl = rg->ctrl.p;
  a.err_base_fin = rg->grid_sell;
  a.base = base;
  a.damping = damping;
  a.tolerance = tolerance;
  a.n_peers = 0;
  a.hot_count = p->hot_count;
  return a;
}

// ---- drivers ---------------------------------------------------------------------------------
static gb_status run_exact(const gb_graph* g, const gb_page_rank_config* cfg, float* d_scores,
                           uint64_t* ran, double* error) {
  cudaStream_t s = g->stream;
  DevBuf<float> out;
  DevBuf<uint64_t> d_ran;
  DevBuf<double> d_err;
  GB_TRY(out.alloc(g->n));
  GB_TRY(d_ran.alloc(1));
  GB_TRY(d_err.alloc(1));
  k_pr_exact<<<1, 32, 0, s>>>(g->in.off.p, g->in.tgt.p, g->out.off.p, g->n, cfg->max_iterations,
                             cfg->tolerance, cfg->damping_factor, d_scores, out.p, d_ran.p, d_err.p);
  GB_CUDA(cudaGetLastError());
  g->timing.kernel_launches += 1;
  GB_CUDA(cudaMemcpyAsync(ran, d_ran.p, 8, cudaMemcpyDeviceToHost, s));
  GB_CUDA(cudaMemcpyAsync(error, d_err.p, 8, cudaMemcpyDeviceToHost, s));
  GB_CUDA(cudaStreamSynchronize(s));
  return GB_OK;
}

static gb_status run_jacobi(const gb_graph* g, const gb_page_rank_config* cfg, float* d_scores,
                            uint64_t* ran, double* error) {
  if (!g->pr_plan) GB_TRY(build_pr_plan(g, &g->pr_plan));
  PrPlan* p = g->pr_plan;
  cudaStream_t s = g->stream;
  const uint32_t n = p->n;
  const float nf = (float)n;
  const float init = 1.0f / nf;                             // page_rank.rs:70
  const float base = (1.0f - cfg->damping_factor) / nf;     // page_rank.rs:71
  const bool profile = profiling_on();

  k_pr_init<<<grid_for(n, 256), 256, 0, s>>>(n, p->n_active, init, base, p->outdeg.p, p->x[0].p, p->x[1].p,
                                            p->scores.p);
  PrRange* rg = &p->all;
  GB_CUDA(cudaMemsetAsync(rg->ctrl.p, 0, 8, s));
  g->timing.kernel_launches += 1;

  PrArgs a = make_args(p, rg, base, cfg->damping_factor, cfg->tolerance);
  a.scores = p->scores.p;

  // max_iterations == 0 never satisfies `iteration == max_iterations` (page_rank.rs:107): the
  // reference then runs until the tolerance is met; we bound that at 100000 sweeps.
  const uint64_t limit = cfg->max_iterations ? cfg->max_iterations : 100000ull;
  const bool can_stop_early = cfg->tolerance > 0.0;
  const uint32_t batch_cap = 64;
  uint64_t done = 0;      // sweeps launched so far
  uint64_t stopped = 0;   // sweep number at which the tolerance was met (0 = not yet)
  double last_err = 0.0;
  size_t ev_used = 0;
  while (done < limit && !stopped) {
    const uint32_t batch = (uint32_t)std::min<uint64_t>(limit - done, can_stop_early ? 8 : batch_cap);
    for (uint32_t b = 0; b < batch; ++b) {
      const uint64_t sweep_no = done + b + 1;
      a.x_cur = p->x[(sweep_no - 1) & 1].p;
      a.x_next = p->x[sweep_no & 1].p;
      a.sweep = b;
      a.sweep_no = (uint32_t)std::min<uint64_t>(sweep_no, 0xFFFFFFFFull);
      a.extra_err = (sweep_no == 1)
                        ? (double)(n - p->n_active) * fabs((double)(base - init))
                        : 0.0;
      cudaEvent_t e0 = nullptr, e1 = nullptr;
      if (profile && ev_used + 2 <= 2 * PR_MAX_PROFILE_EVENTS) {
        while (p->prof_events.size() < ev_used + 2) {
          cudaEvent_t e;
          GB_CUDA(cudaEventCreate(&e));
          p->prof_events.push_back(e);
        }
        e0 = p->prof_events[ev_used];
        e1 = p->prof_events[ev_used + 1];
        ev_used += 2;
        GB_CUDA(cudaEventRecord(e0, s));
      }
      if (rg->grid_seg) k_pr_seg<<<rg->grid_seg, PR_THREADS, p->smem_bytes, s>>>(a);
      if (rg->grid_sell) k_pr_sell<false><<<rg->grid_sell, PR_THREADS, p->smem_bytes, s>>>(a);
      if (e1) GB_CUDA(cudaEventRecord(e1, s));
      k_pr_finish<false><<<rg->grid_fin, PR_FIN_THREADS, 0, s>>>(a);
      g->timing.kernel_launches += 1 + (rg->grid_seg ? 1 : 0) + (rg->grid_sell ? 1 : 0);
      if (sweep_no == 1 && p->n_active < n) {
        // sources without in-edges change exactly once (init/deg -> base/deg): patch the buffer
        // sweep 1 has just finished reading
        k_pr_fill_inactive<<<grid_for(n - p->n_active, 256), 256, 0, s>>>(n, p->n_active, base, p->outdeg.p,
                                                                         p->x[0].p);
        g->timing.kernel_launches += 1;
      }
    }
    GB_CUDA(cudaGetLastError());
    done += batch;
    if (can_stop_early || done >= limit) {
      uint32_t ctrl0 = 0;
      GB_CUDA(cudaMemcpyAsync(&ctrl0, rg->ctrl.p, 4, cudaMemcpyDeviceToHost, s));
      GB_CUDA(cudaStreamSynchronize(s));
      if (ctrl0 != 0) stopped = ctrl0;
      const uint64_t last = stopped ? stopped : done;
      const uint32_t slot = (uint32_t)(last - (done - batch) - 1);
      GB_CUDA(cudaMemcpyAsync(&last_err, rg->err_hist.p + slot, 8, cudaMemcpyDeviceToHost, s));
      GB_CUDA(cudaStreamSynchronize(s));
    }
  }
  *ran = stopped ? stopped : done;
  *error = last_err;
  k_unpermute<<<grid_for(n, 256), 256, 0, s>>>(p->scores.p, p->new_id.p, n, d_scores);
  g->timing.kernel_launches += 1;
  GB_CUDA(cudaGetLastError());
  GB_CUDA(cudaStreamSynchronize(s));
  if (profile) {
    double ms = 0.0;
    for (size_t i = 0; i + 1 < ev_used; i += 2) {
      float t = 0.0f;
      GB_CUDA(cudaEventElapsedTime(&t, p->prof_events[i], p->prof_events[i + 1]));
      ms += t;
    }
    g->timing.hot_kernel_ms = ms;
    g->timing.hot_kernel_launches = ev_used / 2;
  }
  return GB_OK;
}

static gb_status page_rank_impl(const gb_graph* g, const gb_page_rank_config* cfg, float* d_scores,
                                float* h_scores, uint64_t* ran, double* error) {
  GB_REQUIRE(g && cfg && ran && error, "NULL argument");
  if (g->kind != GB_KIND_DIRECTED)
    return fail(GB_ERR_UNSUPPORTED, "page_rank needs a directed graph (page_rank.rs:61)");
  GB_REQUIRE(cfg->mode <= GB_PR_JACOBI, "bad page rank mode %u", cfg->mode);
  GB_REQUIRE(!(cfg->max_iterations == 0 && !(cfg->tolerance > 0.0)),
             "max_iterations == 0 with tolerance <= 0 never terminates (page_rank.rs:107)");
  DeviceGuard guard(g->device);
  std::lock_guard<std::mutex> lock(g->mu);
  cudaStream_t s = g->stream;
  uint32_t mode = cfg->mode;
  if (mode == GB_PR_AUTO) mode = (g->n <= 16384) ? GB_PR_EXACT : GB_PR_JACOBI;
  DevBuf<float> tmp_scores;
  if (!d_scores) {
    GB_TRY(tmp_scores.alloc(g->n));
    d_scores = tmp_scores.p;
  }
  if (mode == GB_PR_JACOBI && !g->pr_plan) GB_TRY(build_pr_plan(g, &g->pr_plan));  // not timed
  g->timing = gb_timing{};
  GB_CUDA(cudaEventRecord(g->ev_begin, s));
  if (mode == GB_PR_EXACT) GB_TRY(run_exact(g, cfg, d_scores, ran, error));
  else GB_TRY(run_jacobi(g, cfg, d_scores, ran, error));
  GB_CUDA(cudaEventRecord(g->ev_end, s));
  if (h_scores) GB_CUDA(cudaMemcpyAsync(h_scores, d_scores, (size_t)g->n * 4, cudaMemcpyDeviceToHost, s));
  GB_CUDA(cudaStreamSynchronize(s));
  float ms = 0.0f;
  GB_CUDA(cudaEventElapsedTime(&ms, g->ev_begin, g->ev_end));
  g->timing.total_ms = ms;
  return GB_OK;
}


// ---- multi-GPU shard (1-D edge-cut by destination range) ----------------------------------------
}  // namespace gb

struct gb_pr_shard {
  const gb_graph* graph = nullptr;
  gb::PrRange range;
};

namespace gb {

static gb_status shard_partition(const gb_graph* g, uint32_t parts, uint64_t row_cost, const double* cuts,
                                 uint32_t* ranges) {
  // greedy_node_map_partition (graph_ops.rs:479-509) over the INTERNAL row order with
  // node_map = in-degree and batch = ceil(m / parts) (in_degree_partition, graph_ops.rs:431-439)
  if (!g->pr_plan) GB_TRY(build_pr_plan(g, &g->pr_plan));
  const PrPlan* p = g->pr_plan;
  std::vector<uint32_t> off((size_t)p->n + 1);
  GB_CUDA(cudaMemcpyAsync(off.data(), p->off.p, off.size() * 4, cudaMemcpyDeviceToHost, g->stream));
  GB_CUDA(cudaStreamSynchronize(g->stream));
  // node_map = in-degree + a constant per-row charge: a row costs its gathers plus ~5 vector accesses,
  // a division and (multi-GPU) one store per peer, so ranks owning millions of 1-edge rows would
  // otherwise be the stragglers.  GB_SHARD_ROW_COST overrides the charge (0 = the plain rule).
  if (const char* e = getenv("GB_SHARD_ROW_COST")) row_cost = (uint64_t)atoll(e);
  const uint64_t total = p->m + row_cost * p->n_active;
  const uint64_t batch = (total + parts - 1) / parts;
  uint32_t count = 0;
  uint64_t acc = 0;
  ranges[0] = 0;
  if (cuts) {
    // explicit cut points (fractions of the total weight, increasing): used by the measured-time
    // rebalancing of the multi-GPU orchestration
    for (uint32_t k = 0; k + 1 < parts; ++k)
      GB_REQUIRE(cuts[k] > 0.0 && cuts[k] < 1.0 && (k == 0 || cuts[k] >= cuts[k - 1]), "bad cut fraction %u", k);
    for (uint32_t v = 0; v < p->n && count < parts - 1; ++v) {
      const uint32_t d = off[v + 1] - off[v];
      acc += d + (d ? row_cost : 0);
      while (count < parts - 1 && (double)acc >= cuts[count] * (double)total) ranges[++count] = v + 1;
    }
    while (count < parts) ranges[++count] = p->n;
  } else {
    for (uint32_t v = 0; v < p->n; ++v) {
      const uint32_t d = off[v + 1] - off[v];
      acc += d + (d ? row_cost : 0);
      if ((count < parts - 1 && acc >= batch) || v == p->n - 1) {
        ranges[++count] = v + 1;
        acc = 0;
      }
    }
  }
  for (uint32_t i = count + 1; i <= parts; ++i) ranges[i] = p->n;
  // boundaries inside the SELL region move to the nearest slice boundary (32 rows)
  for (uint32_t i = 1; i < parts; ++i) {
    uint32_t b = ranges[i];
    if (b > p->n_long && b < p->n_active) {
      b = p->n_long + ((b - p->n_long + 16) / 32) * 32;
      if (b > p->n_active) b = p->n_active;
    }
    if (b < ranges[i - 1]) b = ranges[i - 1];
    ranges[i] = b;
  }
  return GB_OK;
}

}  // namespace gb

extern "C" {

gb_status gb_pr_shard_partition(const gb_graph* g, uint32_t parts, uint32_t row_cost, const double* cuts,
                                uint32_t* ranges) {
  GB_REQUIRE(g && ranges, "NULL argument");
  GB_REQUIRE(parts >= 1, "parts must be >= 1");
  if (g->kind != GB_KIND_DIRECTED) return gb::fail(GB_ERR_UNSUPPORTED, "page rank shards need a directed graph");
  gb::DeviceGuard guard(g->device);
  std::lock_guard<std::mutex> lock(g->mu);
  return gb::shard_partition(g, parts, row_cost, cuts, ranges);
}

gb_status gb_pr_shard_create(const gb_graph* g, uint32_t row_begin, uint32_t row_end, gb_pr_shard** shard) {
  GB_REQUIRE(g && shard, "NULL argument");
  if (g->kind != GB_KIND_DIRECTED) return gb::fail(GB_ERR_UNSUPPORTED, "page rank shards need a directed graph");
  GB_REQUIRE(row_begin <= row_end && row_end <= g->n, "bad row range [%u, %u)", row_begin, row_end);
  gb::DeviceGuard guard(g->device);
  std::lock_guard<std::mutex> lock(g->mu);
  if (!g->pr_plan) GB_TRY(gb::build_pr_plan(g, &g->pr_plan));
  gb_pr_shard* sh = new (std::nothrow) gb_pr_shard();
  if (!sh) return gb::fail(GB_ERR_OOM, "host allocation failed");
  sh->graph = g;
  gb_status st = gb::build_range(g, g->pr_plan, row_begin, row_end, &sh->range);
  if (st != GB_OK) {
    delete sh;
    return st;
  }
  *shard = sh;
  return GB_OK;
}

gb_status gb_pr_shard_free(gb_pr_shard* shard) {
  if (!shard) return GB_OK;
  gb::DeviceGuard guard(shard->graph->device);
  delete shard;
  return GB_OK;
}

gb_status gb_pr_shard_init(const gb_pr_shard* shard, float damping, float* d_x0, float* d_x1,
                           float* d_scores, void* cuda_stream) {
  GB_REQUIRE(shard && d_x0 && d_x1 && d_scores, "NULL argument");
  const gb_graph* g = shard->graph;
  const gb::PrPlan* p = g->pr_plan;
  gb::DeviceGuard guard(g->device);
  cudaStream_t s = (cudaStream_t)cuda_stream;
  const float nf = (float)p->n;
  const float init = 1.0f / nf;
  const float base = (1.0f - damping) / nf;
  // every rank fills the whole initial vector itself (no exchange needed before sweep 1)
  gb::k_pr_init<<<gb::grid_for(p->n, 256), 256, 0, s>>>(p->n, p->n_active, init, base, p->outdeg.p, d_x0, d_x1,
                                                       d_scores);
  GB_CUDA(cudaMemsetAsync(shard->range.ctrl.p, 0, 8, s));
  GB_CUDA(cudaGetLastError());
  return GB_OK;
}

gb_status gb_pr_shard_step(const gb_pr_shard* shard, float damping, uint64_t sweep_no, const float* d_x_cur,
                           float* d_x_next, float* const* d_peer_x_next, uint32_t peer_count,
                           float* d_scores, double* d_error, void* cuda_stream) {
  GB_REQUIRE(shard && d_x_cur && d_x_next && d_scores && d_error, "NULL argument");
  GB_REQUIRE(peer_count <= 7, "at most 7 peers");
  GB_REQUIRE(peer_count == 0 || d_peer_x_next, "peer pointer array is NULL");
  GB_REQUIRE(sweep_no >= 1, "sweep_no is 1-based");
  const gb_graph* g = shard->graph;
  const gb::PrPlan* p = g->pr_plan;
  const gb::PrRange* rg = &shard->range;
  gb::DeviceGuard guard(g->device);
  cudaStream_t s = (cudaStream_t)cuda_stream;
  const float nf = (float)p->n;
  const float init = 1.0f / nf;
  const float base = (1.0f - damping) / nf;
  gb::PrArgs a = gb::make_args(p, rg, base, damping, -1.0 /* the caller owns the stop rule */);
  a.x_cur = d_x_cur;
  a.x_next = d_x_next;
  a.scores = d_scores;
  a.n_peers = peer_count;
  for (uint32_t i = 0; i < peer_count; ++i) a.peer_next[i] = d_peer_x_next[i];
  a.err_hist = d_error;
  a.sweep = 0;
  a.sweep_no = (uint32_t)std::min<uint64_t>(sweep_no, 0xFFFFFFFFull);
  // the closed-form error of the rows without in-edges is contributed once, by the shard owning row 0
  a.extra_err = (sweep_no == 1 && rg->row_begin == 0)
                    ? (double)(p->n - p->n_active) * fabs((double)(base - init))
                    : 0.0;
  if (rg->grid_seg) gb::k_pr_seg<<<rg->grid_seg, gb::PR_THREADS, p->smem_bytes, s>>>(a);
  if (peer_count) {
    if (rg->grid_sell) gb::k_pr_sell<true><<<rg->grid_sell, gb::PR_THREADS, p->smem_bytes, s>>>(a);
    gb::k_pr_finish<true><<<rg->grid_fin, gb::PR_FIN_THREADS, 0, s>>>(a);
  } else {
    if (rg->grid_sell) gb::k_pr_sell<false><<<rg->grid_sell, gb::PR_THREADS, p->smem_bytes, s>>>(a);
    gb::k_pr_finish<false><<<rg->grid_fin, gb::PR_FIN_THREADS, 0, s>>>(a);
  }
  if (sweep_no == 1 && p->n_active < p->n)
    gb::k_pr_fill_inactive<<<gb::grid_for(p->n - p->n_active, 256), 256, 0, s>>>(
        p->n, p->n_active, base, p->outdeg.p, const_cast<float*>(d_x_cur));
  GB_CUDA(cudaGetLastError());
  return GB_OK;
}

gb_status gb_pr_shard_finish(const gb_pr_shard* shard, const float* d_scores_internal, float* d_scores_out,
                             void* cuda_stream) {
  GB_REQUIRE(shard && d_scores_internal && d_scores_out, "NULL argument");
  const gb_graph* g = shard->graph;
  const gb::PrPlan* p = g->pr_plan;
  gb::DeviceGuard guard(g->device);
  gb::k_unpermute<<<gb::grid_for(p->n, 256), 256, 0, (cudaStream_t)cuda_stream>>>(d_scores_internal, p->new_id.p,
                                                                                 p->n, d_scores_out);
  GB_CUDA(cudaGetLastError());
  return GB_OK;
}

gb_status gb_pr_shard_info(const gb_pr_shard* shard, uint32_t* row_begin, uint32_t* row_end,
                           uint32_t* active_rows, uint64_t* edges) {
  GB_REQUIRE(shard, "NULL argument");
  const gb::PrRange* rg = &shard->range;
  if (row_begin) *row_begin = rg->row_begin;
  if (row_end) *row_end = rg->row_end;
  if (active_rows) *active_rows = shard->graph->pr_plan->n_active;
  if (edges) {
    const gb::PrPlan* p = shard->graph->pr_plan;
    uint32_t h[2] = {0, 0};
    gb::DeviceGuard guard(shard->graph->device);
    GB_CUDA(cudaMemcpy(&h[0], p->off.p + rg->row_begin, 4, cudaMemcpyDeviceToHost));
    GB_CUDA(cudaMemcpy(&h[1], p->off.p + rg->row_end, 4, cudaMemcpyDeviceToHost));
    *edges = (uint64_t)h[1] - h[0];
  }
  return GB_OK;
}

}  // extern "C"

namespace gb {
}  // namespace gb

extern "C" {

gb_status gb_page_rank(const gb_graph* graph, const gb_page_rank_config* config, float* scores,
                       uint64_t* ran_iterations, double* error) {
  GB_REQUIRE(scores != nullptr, "scores is NULL");
  return gb::page_rank_impl(graph, config, nullptr, scores, ran_iterations, error);
}

gb_status gb_page_rank_csr_u32(int device, uint32_t n, const uint32_t* in_off, const uint32_t* in_tgt,
                               const uint32_t* out_off, const gb_page_rank_config* config, float* scores,
                               uint64_t* ran_iterations, double* error) {
  GB_REQUIRE(scores != nullptr, "scores is NULL");
  GB_REQUIRE(n > 0, "node_count must be > 0");
  GB_REQUIRE(in_off && out_off, "offset arrays are NULL");
  GB_REQUIRE(in_off[n] == out_off[n], "in and out offsets disagree on the edge count");
  gb_graph* g = nullptr;
  GB_TRY(gb::new_graph(device, GB_KIND_DIRECTED, n, &g));
  gb_status st = gb::upload_host_csr(g->stream, n, in_off, in_tgt, nullptr, &g->in, "in");
  if (st == GB_OK) st = gb::upload_host_csr(g->stream, n, out_off, nullptr, nullptr, &g->out, "out");
  if (st == GB_OK) st = gb::page_rank_impl(g, config, nullptr, scores, ran_iterations, error);
  gb_graph_free(g);
  return st;
}

gb_status gb_page_rank_device(const gb_graph* graph, const gb_page_rank_config* config, float* d_scores,
                              uint64_t* ran_iterations, double* error) {
  GB_REQUIRE(d_scores != nullptr, "d_scores is NULL");
  return gb::page_rank_impl(graph, config, d_scores, nullptr, ran_iterations, error);
}

}  // extern "C"
