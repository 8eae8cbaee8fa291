// pagerank.cu — PageRank over the device in-CSR.
//
// Replaces crates/algos/src/page_rank.rs:58-168 (`page_rank`, `page_rank_iteration`).
//
// Two schedules (gb_pr_mode, include/graph_b200.h):
//   EXACT  — the reference's sweep as one thread executes it: in place, CSR-order f32 sums, separate
//            multiply/add (no FMA), IEEE division.  One warp walks the vertices in id order; lanes
//            only parallelise the gather loads, the additions stay sequential.  Bit-exact with the
//            reference wherever the reference is deterministic (n <= 16384 = one chunk).
//   JACOBI — the throughput path.  Vertices are renumbered internally (rows with in-edges first,
//            then by out-degree descending, so the hot part of the gathered vector is contiguous and
//            each row's sources are sorted hot-first); the merged sequence "row 0 edges, row 0 end,
//            row 1 edges, row 1 end, ..." is cut into equal chunks of PR_CHUNK items (merge-path), one
//            warp per chunk.  A warp streams its slice of the target array with 128-bit loads,
//            gathers out_scores, stages the values in shared memory and reduces every row that ends
//            in its chunk; rows that straddle chunks leave a partial sum (carry) that a small fix-up
//            kernel combines in chunk order, so the result is deterministic.  Vertices without
//            in-edges are constant after the first sweep and are skipped from then on.
//
// Algorithmic bytes per sweep: 4m (targets) + 4(n+1) (offsets) + 5*4n (out_scores read+write,
// scores read+write, out-degree read) = 4m + 24n + 4  (BASELINE.md §3).
#include <cub/cub.cuh>
#include <thrust/iterator/counting_iterator.h>

#include <algorithm>

#include "common.cuh"

namespace gb {

constexpr int PR_CHUNK = 252;      // merge-path items per warp task (edges + PR_ROW_COST per row end)
constexpr int PR_ROW_COST = 8;     // items charged per row end: a chunk never ends more than 32 rows
constexpr int PR_SLOTS = 256;      // register slots of a chunk: 8 consecutive edges per lane
constexpr int PR_WARPS = 32;       // warps per CTA of the sweep kernel: one persistent CTA per SM
constexpr int PR_THREADS = PR_WARPS * 32;
constexpr int PR_HOT = 52 * 1024;  // out_scores entries mirrored in shared memory (208 KB)
constexpr int PR_WARP_SMEM = PR_SLOTS + 36 * 4;  // per warp: 256 head bytes + 36 row sums
constexpr int PR_FIX_THREADS = 256;
constexpr uint32_t PR_MAX_PROFILE_EVENTS = 256;

// the merge-path chunking of a contiguous range of internal rows (the whole graph on one GPU, or
// one rank's shard of the 1-D edge-cut)
struct PrRange {
  uint32_t row_begin = 0, row_end = 0;  // internal rows [row_begin, row_end), clipped to active rows
  uint64_t item_base = 0;               // off[row_begin] + row_begin
  uint32_t num_chunks = 0;
  uint32_t num_fix = 0;
  unsigned grid_pull = 1, grid_fix = 1;
  DevBuf<uint2> coord;       // merge-path (row, edge) start of each chunk [num_chunks+1]
  DevBuf<uint32_t> fix;      // chunks whose first row started in an earlier chunk [num_fix]
  DevBuf<float> carry_tail;  // per chunk: partial sum of the row continuing into the next chunk
  DevBuf<float> head_part;   // per chunk: partial sum of a first row continued from earlier chunks
  DevBuf<double> block_err;  // per CTA error partials (pull CTAs, then fix CTAs)
  DevBuf<double> err_hist;   // error of each sweep of the current batch
  DevBuf<uint32_t> ctrl;     // [0] = done flag (sweep number at which tolerance was met), [1] = ticket
  uint64_t bytes() const {
    return coord.bytes() + fix.bytes() + carry_tail.bytes() + head_part.bytes() + block_err.bytes() +
           err_hist.bytes() + ctrl.bytes();
  }
};

struct PrPlan {
  uint32_t n = 0;
  uint32_t n_active = 0;  // rows with in-degree > 0 (renumbered to [0, n_active))
  uint64_t m = 0;
  DevBuf<uint32_t> new_id;   // old id -> internal id
  DevBuf<uint32_t> off;      // internal in-CSR offsets [n+1]
  DevBuf<uint32_t> tgt;      // internal in-CSR targets [m] (+8 slack)
  DevBuf<uint32_t> outdeg;   // out-degree by internal id [n]
  DevBuf<float> x[2];        // out_scores ping-pong [n]
  DevBuf<float> scores;      // ranks by internal id [n]
  PrRange all;               // chunking of every active row (single-GPU path)
  uint32_t hot_count = 0;    // entries of out_scores mirrored in shared memory by the sweep kernel
  size_t smem_bytes = 0;     // dynamic shared memory of the sweep kernel
  std::vector<cudaEvent_t> prof_events;
  uint64_t bytes() const {
    return new_id.bytes() + off.bytes() + tgt.bytes() + outdeg.bytes() + x[0].bytes() + x[1].bytes() +
           scores.bytes() + all.bytes();
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

// the per-vertex update of page_rank.rs:148-158 with the reference's rounding sequence
struct PrArgs;
template <bool PEERS>
__device__ __forceinline__ double pr_finalize(uint32_t r, float sum, const PrArgs& a);

// ---- plan construction kernels ---------------------------------------------------------------
__global__ void k_perm_keys(const uint32_t* __restrict__ in_off, const uint32_t* __restrict__ out_off,
                            uint32_t n, uint64_t* __restrict__ keys, uint32_t* __restrict__ ids) {
  for (uint32_t v = blockIdx.x * blockDim.x + threadIdx.x; v < n; v += gridDim.x * blockDim.x) {
    uint32_t indeg = in_off[v + 1] - in_off[v];
    uint32_t outdeg = out_off[v + 1] - out_off[v];
    keys[v] = ((uint64_t)(indeg == 0) << 32) | (uint32_t)(~outdeg);  // active first, hot first
    ids[v] = v;
  }
}
__global__ void k_perm_scatter(const uint32_t* __restrict__ sorted_ids, const uint32_t* __restrict__ out_off,
                               uint32_t n, uint32_t* __restrict__ new_id, uint32_t* __restrict__ outdeg) {
  for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
    uint32_t v = sorted_ids[r];
    new_id[v] = r;
    outdeg[r] = out_off[v + 1] - out_off[v];
  }
}
// one warp per original row: keys of the renumbered in-CSR
__global__ void k_perm_edge_keys(const uint32_t* __restrict__ in_off, const uint32_t* __restrict__ in_tgt,
                                 const uint32_t* __restrict__ new_id, uint32_t n, uint32_t bits,
                                 uint64_t* __restrict__ keys) {
  uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t v = warp; v < n; v += nwarps) {
    uint32_t b = in_off[v], e = in_off[v + 1];
    if (b == e) continue;
    uint64_t hi = (uint64_t)new_id[v] << bits;
    for (uint32_t i = b + lane; i < e; i += 32) keys[i] = hi | new_id[in_tgt[i]];
  }
}
__global__ void k_unpack_low(const uint64_t* __restrict__ keys, uint64_t count, uint32_t bits,
                             uint32_t* __restrict__ tgt) {
  uint64_t mask = (1ull << bits) - 1ull;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < count;
       i += (uint64_t)gridDim.x * blockDim.x)
    tgt[i] = (uint32_t)(keys[i] & mask);
}
__global__ void k_mark_ends_key(const uint64_t* __restrict__ keys, uint64_t count, uint32_t bits,
                                uint32_t* __restrict__ marks) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < count;
       i += (uint64_t)gridDim.x * blockDim.x) {
    uint32_t r = (uint32_t)(keys[i] >> bits);
    if (i + 1 == count || (uint32_t)(keys[i + 1] >> bits) != r) marks[r + 1] = (uint32_t)(i + 1);
  }
}
// merge-path split: chunk k starts at diagonal k*PR_CHUNK of (row ends) x (edges)
__global__ void k_merge_coords(const uint32_t* __restrict__ off, uint32_t row_begin, uint32_t row_end,
                               uint64_t item_base, uint64_t items, uint32_t num_chunks,
                               uint2* __restrict__ coord) {
  for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k <= num_chunks; k += gridDim.x * blockDim.x) {
    uint64_t diag = (uint64_t)k * PR_CHUNK;
    if (diag > items) diag = items;
    diag += item_base;  // absolute position in the merged sequence: row r's edges, then PR_ROW_COST slots
    // r0 = first row whose end marker (first slot at off[r+1] + COST*r) is not before the diagonal
    uint32_t lo = row_begin, hi = row_end;
    while (lo < hi) {
      uint32_t mid = lo + (hi - lo) / 2;
      if ((uint64_t)off[mid + 1] + (uint64_t)PR_ROW_COST * mid < diag) lo = mid + 1; else hi = mid;
    }
    // edges before the diagonal: all edges of rows < r0 plus the part of row r0 in front of it
    uint32_t e;
    if (lo >= row_end) {
      e = off[row_end];
    } else {
      const int64_t want = (int64_t)diag - (int64_t)PR_ROW_COST * lo;
      const int64_t b = off[lo], t = off[lo + 1];
      e = (uint32_t)(want < b ? b : (want > t ? t : want));
    }
    coord[k] = make_uint2(lo, e);
  }
}
__global__ void k_fix_flags(const uint32_t* __restrict__ off, const uint2* __restrict__ coord,
                            uint32_t num_chunks, uint8_t* __restrict__ flags) {
  for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < num_chunks; k += gridDim.x * blockDim.x) {
    uint2 c0 = coord[k], c1 = coord[k + 1];
    flags[k] = (c0.x < c1.x && off[c0.x] < c0.y) ? 1 : 0;
  }
}

// ---- sweep kernels (JACOBI) ------------------------------------------------------------------
struct PrArgs {
  const uint32_t* off;
  const uint32_t* tgt;
  const uint32_t* outdeg;
  const uint2* coord;
  const uint32_t* fix;
  const float* x_cur;
  float* x_next;
  float* peer_next[7];  // peer-mapped copies of x_next (fused allgather over NVLink); n_peers used
  uint32_t n_peers;
  uint32_t hot_count;   // entries of x_cur mirrored in shared memory (multiple of 4)
  uint64_t item_base;
  float* scores;
  float* carry_tail;
  float* head_part;
  double* block_err;
  double* err_hist;
  uint32_t* ctrl;
  uint32_t row_end, num_chunks, num_fix;  // row_end: one past the last row of this range
  unsigned grid_pull;
  float base, damping;
  double tolerance;
  double extra_err;   // closed-form error of the skipped zero-in-degree rows (first sweep only)
  uint32_t sweep;     // index inside the current batch
  uint32_t sweep_no;  // 1-based global sweep number
};

template <bool PEERS>
__device__ __forceinline__ double pr_finalize(uint32_t r, float sum, const PrArgs& a) {
  const float old = a.scores[r];
  const float nw = __fadd_rn(a.base, __fmul_rn(a.damping, sum));
  a.scores[r] = nw;
  const float xo = __fdiv_rn(nw, (float)a.outdeg[r]);
  a.x_next[r] = xo;
  // fused allgather: the finished out_score also goes straight into every peer's next vector
  if (PEERS)
    for (uint32_t p = 0; p < a.n_peers; ++p) a.peer_next[p][r] = xo;
  return fabs((double)__fsub_rn(nw, old));
}

// The sweep: one persistent CTA per SM.  The first PR_HOT entries of out_scores — the most gathered
// sources, contiguous thanks to the out-degree ordering — are mirrored in shared memory once per
// sweep; gathers of those ids are shared-memory loads (bank-limited, ~10 per clock per SM) instead
// of divergent global loads (L1TEX accepts ~0.57 sectors per clock per SM: the measured ceiling of
// the first version of this kernel).  Targets are read with lane-consecutive 32-bit loads so that
// one gather instruction covers 32 consecutive entries of a sorted row and coalesces wherever a
// row's sources are dense.
struct RowMeta {
  uint32_t os, oe, deg;
  float old;
};
// offsets / out-degree / old score of the rows ending in a chunk (one row per lane, at most 32)
__device__ __forceinline__ RowMeta pr_load_meta(const PrArgs& a, uint32_t r0, uint32_t r1, uint32_t lane) {
  RowMeta m{0u, 0u, 1u, 0.0f};
  const uint32_t r = r0 + lane;
  if (r < r1) {
    m.os = a.off[r];
    m.oe = a.off[r + 1];
    m.deg = a.outdeg[r];
    m.old = a.scores[r];
  }
  return m;
}
// 8 consecutive targets per lane (two aligned 128-bit loads); slots outside [e0, e1) become ~0
__device__ __forceinline__ void pr_load_targets(const uint32_t* __restrict__ tgt, uint32_t e0, uint32_t e1,
                                                uint32_t lane, uint32_t (&t)[8]) {
  const uint32_t i0 = (e0 & ~3u) + 8 * lane;
  uint4 ta = make_uint4(~0u, ~0u, ~0u, ~0u), tb = ta;
  if (i0 < e1) ta = ld_stream_u4(tgt + i0);
  if (i0 + 4 < e1) tb = ld_stream_u4(tgt + i0 + 4);
  t[0] = ta.x; t[1] = ta.y; t[2] = ta.z; t[3] = ta.w;
  t[4] = tb.x; t[5] = tb.y; t[6] = tb.z; t[7] = tb.w;
#pragma unroll
  for (int j = 0; j < 8; ++j)
    if (i0 + j < e0 || i0 + j >= e1) t[j] = ~0u;
}
__device__ __forceinline__ void pr_gather(const float* __restrict__ x, const float* hot, uint32_t hot_n,
                                          const uint32_t (&t)[8], float (&v)[8]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const uint32_t tj = t[j];
    float val = 0.0f;
    if (tj < hot_n) val = hot[tj];
    else if (tj != ~0u) val = __ldg(x + tj);
    v[j] = val;
  }
}

template <bool PEERS>
__device__ __forceinline__ double pr_update(uint32_t r, float sum, float old, uint32_t deg, const PrArgs& a) {
  const float nw = __fadd_rn(a.base, __fmul_rn(a.damping, sum));
  a.scores[r] = nw;
  const float xo = __fdiv_rn(nw, (float)deg);
  a.x_next[r] = xo;
  if (PEERS)
    for (uint32_t p = 0; p < a.n_peers; ++p) a.peer_next[p][r] = xo;
  return fabs((double)__fsub_rn(nw, old));
}

// The sweep: one persistent CTA per SM.
//  * The first PR_HOT entries of out_scores — the most gathered sources, contiguous thanks to the
//    out-degree ordering — are mirrored in shared memory once per sweep; gathers of those ids are
//    shared-memory loads instead of divergent global loads (L1TEX accepts only ~0.57 divergent
//    sectors per clock per SM: the measured ceiling of the first version of this kernel).
//  * A chunk is <= 252 consecutive edges and <= 32 row ends.  Each lane owns 8 consecutive edges
//    (two 128-bit loads of the target stream), gathers them into registers, sums its own run
//    between row boundaries and one warp-level segmented scan joins the runs across lanes; the row
//    totals meet their rows (one lane per row, metadata prefetched) through 33 floats of smem.
//  * Two-deep software pipeline: while chunk k is reduced, the gathers of chunk k+1, the targets of
//    chunk k+2 and the coordinates of chunk k+3 are already in flight.
template <bool PEERS>
__global__ void __launch_bounds__(PR_THREADS, 1) k_pr_pull(const PrArgs a) {
  extern __shared__ __align__(16) float smem[];
  float* hot = smem;  // [hot_count]
  unsigned char* warp_base = reinterpret_cast<unsigned char*>(smem + a.hot_count);
  __shared__ double warp_err[PR_WARPS];
  if (a.ctrl[0] != 0) return;  // tolerance already met by an earlier sweep of this batch
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned char* heads = warp_base + warp * PR_WARP_SMEM;            // [256] row id + 1 at the slot a row starts
  float* rs = reinterpret_cast<float*>(heads + PR_SLOTS);            // [36] row totals of this chunk
  const float* __restrict__ x = a.x_cur;
  const uint32_t hot_n = a.hot_count;
  for (uint32_t i = threadIdx.x * 4; i < hot_n; i += PR_THREADS * 4)
    *reinterpret_cast<float4*>(hot + i) = __ldg(reinterpret_cast<const float4*>(x + i));
  __syncthreads();
  double err = 0.0;

  const uint32_t stride = gridDim.x * PR_WARPS;
  const uint32_t K = a.num_chunks;
  uint32_t k = blockIdx.x * PR_WARPS + warp;
  uint2 c0 = make_uint2(0, 0), c1 = c0, n0 = c0, n1 = c0, m0 = c0, m1 = c0;
  uint32_t t[8];
  float v[8];
  RowMeta cur{0u, 0u, 1u, 0.0f}, nxt = cur;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    t[j] = ~0u;
    v[j] = 0.0f;
  }
  if (k < K) {
    c0 = a.coord[k];
    c1 = a.coord[k + 1];
    if (k + stride < K) {
      n0 = a.coord[k + stride];
      n1 = a.coord[k + stride + 1];
    }
    if (k + 2 * stride < K) {
      m0 = a.coord[k + 2 * stride];
      m1 = a.coord[k + 2 * stride + 1];
    }
    pr_load_targets(a.tgt, c0.y, c1.y, lane, t);
    cur = pr_load_meta(a, c0.x, c1.x, lane);
    pr_gather(x, hot, hot_n, t, v);
    if (k + stride < K) pr_load_targets(a.tgt, n0.y, n1.y, lane, t);
  }
  while (k < K) {
    const uint32_t r0 = c0.x, e0 = c0.y, r1 = c1.x, e1 = c1.y;
    const uint32_t a0 = e0 & ~3u;
    const uint32_t nr = r1 - r0;  // rows ending in this chunk (<= 32)
    const uint32_t kn = k + stride, knn = kn + stride;

    // ---- heads: where does each row of this chunk start among the 256 slots? -----------------
    uint32_t seg_s = 0, seg_e = 0;
    bool tail_exists = false;
    if (nr) {
      *reinterpret_cast<uint2*>(heads + 8 * lane) = make_uint2(0u, 0u);
      __syncwarp();
      if (lane < nr) {
        seg_s = cur.os > e0 ? cur.os : e0;
        seg_e = cur.oe;
        if (seg_e > seg_s) heads[seg_s - a0] = (unsigned char)(lane + 1);
      }
      // the row that continues into the next chunk starts where the last ending row stops
      const uint32_t last_oe = __shfl_sync(0xFFFFFFFFu, cur.oe, nr - 1);
      tail_exists = last_oe < e1;
      if (lane == 0 && tail_exists) heads[last_oe - a0] = (unsigned char)(nr + 1);
      __syncwarp();
    }

    // ---- lane-local runs between row starts -----------------------------------------------------
    float run = 0.0f, head_sum = 0.0f;
    uint32_t first_f = 0, prev_f = 0;
    if (nr) {
      const uint2 hb = *reinterpret_cast<const uint2*>(heads + 8 * lane);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint32_t f = ((j < 4 ? hb.x : hb.y) >> (8 * (j & 3))) & 0xFFu;
        if (f) {
          if (prev_f == 0) {
            head_sum = run;  // belongs to the run entering this lane
            first_f = f;
          } else {
            rs[prev_f - 1] = run;  // a row that starts and ends inside this lane
          }
          prev_f = f;
          run = 0.0f;
        }
        run += v[j];
      }
    } else {
      run = ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]));
    }

    // ---- v is consumed: issue the next chunk's gathers and the loads behind them ---------------
    uint2 q0 = make_uint2(0, 0), q1 = q0;
    if (kn < K) {
      pr_gather(x, hot, hot_n, t, v);
      nxt = pr_load_meta(a, n0.x, n1.x, lane);
      if (knn < K) {
        pr_load_targets(a.tgt, m0.y, m1.y, lane, t);
        if (knn + stride < K) {
          q0 = a.coord[knn + stride];
          q1 = a.coord[knn + stride + 1];
        }
      }
    }

    if (nr == 0) {
      // the whole chunk lies inside one (long) row
      const float acc = warp_sum(run);
      if (lane == 0) a.carry_tail[k] = acc;
    } else {
      // ---- segmented inclusive scan of the lane runs (a lane with a row start blocks the carry) --
      const unsigned flagged = __ballot_sync(0xFFFFFFFFu, prev_f != 0);
      float val = run;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const float up = __shfl_up_sync(0xFFFFFFFFu, val, d);
        // lanes (lane-d, lane] must hold no row start for the carry to pass
        const unsigned window = (lane >= (uint32_t)d) ? ((flagged >> (lane - d + 1)) & ((1u << d) - 1u)) : 1u;
        if (window == 0) val += up;
      }
      float carry_in = __shfl_up_sync(0xFFFFFFFFu, val, 1);
      if (lane == 0) carry_in = 0.0f;
      // the run that entered this lane ends at its first row start: it is row first_f - 2
      if (first_f >= 2) rs[first_f - 2] = carry_in + head_sum;
      // the last run of the chunk: total sits in lane 31, its row id in the last flagged lane
      const int last_lane = 31 - __clz(flagged);  // flagged != 0: row 0 or row 1 starts in this chunk
      const uint32_t last_f = __shfl_sync(0xFFFFFFFFu, prev_f, last_lane);
      if (lane == 31 && last_f) rs[last_f - 1] = val;
      __syncwarp();
      // ---- one lane per row: finish it ----------------------------------------------------------
      if (lane < nr) {
        const float sum = (seg_e > seg_s) ? rs[lane] : 0.0f;
        if (cur.os < e0) a.head_part[k] = sum;  // row began in an earlier chunk: fix-up kernel finishes it
        else err += pr_update<PEERS>(r0 + lane, sum, cur.old, cur.deg, a);
      }
      if (lane == 0) a.carry_tail[k] = tail_exists ? rs[nr] : 0.0f;
      __syncwarp();
    }

    k = kn;
    c0 = n0;
    c1 = n1;
    n0 = m0;
    n1 = m1;
    m0 = q0;
    m1 = q1;
    cur = nxt;
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

// combines the partial sums of rows that straddle chunks (one warp per such row) and, in the last
// CTA to finish, reduces the error of the sweep in a fixed order and evaluates the stop rule of
// page_rank.rs:107.
template <bool PEERS>
__global__ void __launch_bounds__(PR_FIX_THREADS) k_pr_fix(const PrArgs a) {
  constexpr int FIX_WARPS = PR_FIX_THREADS / 32;
  __shared__ double warp_err[FIX_WARPS];
  __shared__ bool is_last;
  if (a.ctrl[0] != 0) return;
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double err = 0.0;
  const uint32_t nthreads = gridDim.x * PR_FIX_THREADS;
  // one lane per straddling row; rows spanning many chunks (hubs) are summed by the whole warp
  for (uint32_t ib = blockIdx.x * PR_FIX_THREADS + warp * 32; ib < a.num_fix; ib += nthreads) {
    const uint32_t i = ib + lane;
    const bool live = i < a.num_fix;
    uint32_t k = 0, r0 = 0, j0 = 0;
    if (live) {
      k = a.fix[i];
      r0 = a.coord[k].x;
      j0 = (uint32_t)(((uint64_t)a.off[r0] + (uint64_t)PR_ROW_COST * r0 - a.item_base) / PR_CHUNK);  // chunk holding the row's first edge
    }
    const uint32_t span = k - j0;
    const bool is_long = live && span > 8;
    float p = 0.0f;
    if (live && !is_long)
      for (uint32_t j = j0; j < k; ++j) p += a.carry_tail[j];
    unsigned long_mask = __ballot_sync(0xFFFFFFFFu, is_long);
    while (long_mask) {
      const int owner = __ffs(long_mask) - 1;
      long_mask &= long_mask - 1;
      const uint32_t oj = __shfl_sync(0xFFFFFFFFu, j0, owner);
      const uint32_t ok = __shfl_sync(0xFFFFFFFFu, k, owner);
      float q = 0.0f;
      for (uint32_t j = oj + lane; j < ok; j += 32) q += a.carry_tail[j];
      q = warp_sum(q);
      if ((int)lane == owner) p = q;
    }
    if (live) err += pr_finalize<PEERS>(r0, p + a.head_part[k], a);
  }
  err = warp_sum(err);
  if (lane == 0) warp_err[warp] = err;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < FIX_WARPS; ++w) t += warp_err[w];
    a.block_err[a.grid_pull + blockIdx.x] = t;
    __threadfence();
    unsigned ticket = atomicAdd(&a.ctrl[1], 1u);
    is_last = (ticket == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  // fixed-order reduction of all CTA partials (deterministic error)
  const uint32_t total = a.grid_pull + gridDim.x;
  double t = 0.0;
  for (uint32_t i = threadIdx.x; i < total; i += PR_FIX_THREADS) t += ((volatile double*)a.block_err)[i];
  t = warp_sum(t);
  if (lane == 0) warp_err[warp] = t;
  __syncthreads();
  if (threadIdx.x == 0) {
    double e = a.extra_err;
#pragma unroll
    for (int w = 0; w < FIX_WARPS; ++w) e += warp_err[w];
    a.err_hist[a.sweep] = e;
    a.ctrl[1] = 0;
    if (e < a.tolerance) a.ctrl[0] = a.sweep_no;
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
  uint32_t h_off[2] = {0, 0};
  GB_CUDA(cudaMemcpyAsync(&h_off[0], p->off.p + row_begin, 4, cudaMemcpyDeviceToHost, s));
  GB_CUDA(cudaMemcpyAsync(&h_off[1], p->off.p + row_end, 4, cudaMemcpyDeviceToHost, s));
  GB_CUDA(cudaStreamSynchronize(s));
  r->item_base = (uint64_t)h_off[0] + (uint64_t)PR_ROW_COST * row_begin;
  const uint64_t items = (uint64_t)(h_off[1] - h_off[0]) + (uint64_t)PR_ROW_COST * (row_end - row_begin);
  const uint64_t nchunks = (items + PR_CHUNK - 1) / PR_CHUNK;
  GB_REQUIRE(nchunks < 0xFFFFFFF0ull, "too many chunks");
  r->num_chunks = (uint32_t)nchunks;
  GB_TRY(r->coord.alloc((size_t)r->num_chunks + 1));
  k_merge_coords<<<grid_for((uint64_t)r->num_chunks + 1, 256), 256, 0, s>>>(
      p->off.p, row_begin, row_end, r->item_base, items, r->num_chunks, r->coord.p);
  GB_TRY(r->carry_tail.alloc(std::max<size_t>(r->num_chunks, 1)));
  GB_TRY(r->head_part.alloc(std::max<size_t>(r->num_chunks, 1)));
  r->num_fix = 0;
  GB_TRY(r->fix.alloc(std::max<size_t>(r->num_chunks, 1)));
  if (r->num_chunks) {
    DevBuf<uint8_t> flags;
    DevBuf<uint32_t> d_num;
    GB_TRY(flags.alloc(r->num_chunks));
    GB_TRY(d_num.alloc(1));
    k_fix_flags<<<grid_for(r->num_chunks, 256), 256, 0, s>>>(p->off.p, r->coord.p, r->num_chunks, flags.p);
    thrust::counting_iterator<uint32_t> iota(0);
    size_t tb = 0;
    GB_CUDA(cub::DeviceSelect::Flagged(nullptr, tb, iota, flags.p, r->fix.p, d_num.p, (int)r->num_chunks, s));
    DevBuf<uint8_t> tmp;
    GB_TRY(tmp.alloc(tb));
    GB_CUDA(cub::DeviceSelect::Flagged(tmp.p, tb, iota, flags.p, r->fix.p, d_num.p, (int)r->num_chunks, s));
    GB_CUDA(cudaMemcpyAsync(&r->num_fix, d_num.p, 4, cudaMemcpyDeviceToHost, s));
    GB_CUDA(cudaStreamSynchronize(s));
  }
  int dev_sms = 148;
  GB_CUDA(cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, g->device));
  const uint64_t want = ((uint64_t)r->num_chunks + PR_WARPS - 1) / PR_WARPS;
  r->grid_pull = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(want, (uint64_t)dev_sms));  // 1 CTA / SM
  const uint64_t want_fix = ((uint64_t)r->num_fix + PR_FIX_THREADS - 1) / PR_FIX_THREADS;
  r->grid_fix = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(want_fix, (uint64_t)dev_sms * 8));
  GB_TRY(r->block_err.alloc((size_t)r->grid_pull + r->grid_fix));
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
    uint32_t bits = 1;
    while (bits < 32 && (1ull << bits) < n) ++bits;
    // 1. permutation: rows with in-edges first, then out-degree descending, then id
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
      GB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, kb, vb, (int)n, 0, 33, s));
      DevBuf<uint8_t> tmp;
      GB_TRY(tmp.alloc(tb));
      GB_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tb, kb, vb, (int)n, 0, 33, s));
      GB_TRY(p->new_id.alloc(n));
      GB_TRY(p->outdeg.alloc(n));
      k_perm_scatter<<<grid_for(n, 256), 256, 0, s>>>(vb.Current(), g->out.off.p, n, p->new_id.p,
                                                     p->outdeg.p);
      GB_CUDA(cudaGetLastError());
      GB_CUDA(cudaStreamSynchronize(s));
    }
    // 2. renumbered in-CSR (rows sorted by internal source id: hot sources first)
    GB_TRY(p->off.alloc((size_t)n + 1));
    GB_TRY(p->tgt.alloc(m, 8));
    GB_CUDA(cudaMemsetAsync(p->off.p, 0, ((size_t)n + 1) * 4, s));
    GB_CUDA(cudaMemsetAsync(p->tgt.p + m, 0, 8 * 4, s));
    if (m) {
      DevBuf<uint64_t> keys, keys_alt;
      GB_TRY(keys.alloc(m));
      GB_TRY(keys_alt.alloc(m));
      k_perm_edge_keys<<<grid_for((uint64_t)n * 32, 256), 256, 0, s>>>(g->in.off.p, g->in.tgt.p,
                                                                      p->new_id.p, n, bits, keys.p);
      cub::DoubleBuffer<uint64_t> kb(keys.p, keys_alt.p);
      size_t tb = 0;
      GB_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tb, kb, m, 0, (int)(2 * bits), s));
      DevBuf<uint8_t> tmp;
      GB_TRY(tmp.alloc(tb));
      GB_CUDA(cub::DeviceRadixSort::SortKeys(tmp.p, tb, kb, m, 0, (int)(2 * bits), s));
      k_unpack_low<<<grid_for(m, 256), 256, 0, s>>>(kb.Current(), m, bits, p->tgt.p);
      k_mark_ends_key<<<grid_for(m, 256), 256, 0, s>>>(kb.Current(), m, bits, p->off.p);
      GB_CUDA(cudaGetLastError());
      GB_CUDA(cudaStreamSynchronize(s));
    }
    {
      size_t tb = 0;
      GB_CUDA(cub::DeviceScan::InclusiveScan(nullptr, tb, p->off.p, p->off.p, cub::Max(), (int64_t)n + 1, s));
      DevBuf<uint8_t> tmp;
      GB_TRY(tmp.alloc(tb));
      GB_CUDA(cub::DeviceScan::InclusiveScan(tmp.p, tb, p->off.p, p->off.p, cub::Max(), (int64_t)n + 1, s));
      GB_CUDA(cudaStreamSynchronize(s));
    }
    // 3. n_active = first row whose offset equals m (rows are ordered active-first)
    {
      std::vector<uint32_t> probe(1);
      // binary search on device memory through small copies (log2(n) 4-byte reads)
      uint32_t lo = 0, hi = n;
      while (lo < hi) {
        uint32_t mid = lo + (hi - lo) / 2;
        GB_CUDA(cudaMemcpyAsync(probe.data(), p->off.p + mid, 4, cudaMemcpyDeviceToHost, s));
        GB_CUDA(cudaStreamSynchronize(s));
        if (probe[0] < m) lo = mid + 1; else hi = mid;
      }
      // lo = first row r with off[r] >= m  => rows [lo, n) are empty, but row lo-1 may also be
      // empty only if m == 0
      p->n_active = (m == 0) ? 0 : lo;
    }
    p->hot_count = std::min<uint32_t>((uint32_t)PR_HOT, n & ~3u);
    p->smem_bytes = (size_t)p->hot_count * sizeof(float) + (size_t)PR_WARPS * PR_WARP_SMEM;
    GB_CUDA(cudaFuncSetAttribute(k_pr_pull<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem_bytes));
    GB_CUDA(cudaFuncSetAttribute(k_pr_pull<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem_bytes));
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
  a.off = p->off.p;
  a.tgt = p->tgt.p;
  a.outdeg = p->outdeg.p;
  a.coord = rg->coord.p;
  a.fix = rg->fix.p;
  a.carry_tail = rg->carry_tail.p;
  a.head_part = rg->head_part.p;
  a.block_err = rg->block_err.p;
  a.err_hist = rg->err_hist.p;
  a.ctrl = rg->ctrl.p;
  a.row_end = rg->row_end;
  a.item_base = rg->item_base;
  a.num_chunks = rg->num_chunks;
  a.num_fix = rg->num_fix;
  a.grid_pull = rg->grid_pull;
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
      if (rg->num_chunks) k_pr_pull<false><<<rg->grid_pull, PR_THREADS, p->smem_bytes, s>>>(a);
      if (e1) GB_CUDA(cudaEventRecord(e1, s));
      k_pr_fix<false><<<rg->grid_fix, PR_FIX_THREADS, 0, s>>>(a);
      g->timing.kernel_launches += rg->num_chunks ? 2 : 1;
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

static gb_status shard_partition(const gb_graph* g, uint32_t parts, uint32_t* ranges) {
  // greedy_node_map_partition (graph_ops.rs:479-509) over the INTERNAL row order with
  // node_map = in-degree and batch = ceil(m / parts) (in_degree_partition, graph_ops.rs:431-439)
  if (!g->pr_plan) GB_TRY(build_pr_plan(g, &g->pr_plan));
  const PrPlan* p = g->pr_plan;
  std::vector<uint32_t> off((size_t)p->n + 1);
  GB_CUDA(cudaMemcpyAsync(off.data(), p->off.p, off.size() * 4, cudaMemcpyDeviceToHost, g->stream));
  GB_CUDA(cudaStreamSynchronize(g->stream));
  const uint64_t batch = (p->m + parts - 1) / parts;
  uint32_t count = 0;
  uint64_t acc = 0;
  ranges[0] = 0;
  for (uint32_t v = 0; v < p->n; ++v) {
    acc += off[v + 1] - off[v];
    if ((count < parts - 1 && acc >= batch) || v == p->n - 1) {
      ranges[++count] = v + 1;
      acc = 0;
    }
  }
  for (uint32_t i = count + 1; i <= parts; ++i) ranges[i] = p->n;
  return GB_OK;
}

}  // namespace gb

extern "C" {

gb_status gb_pr_shard_partition(const gb_graph* g, uint32_t parts, uint32_t* ranges) {
  GB_REQUIRE(g && ranges, "NULL argument");
  GB_REQUIRE(parts >= 1, "parts must be >= 1");
  if (g->kind != GB_KIND_DIRECTED) return gb::fail(GB_ERR_UNSUPPORTED, "page rank shards need a directed graph");
  gb::DeviceGuard guard(g->device);
  std::lock_guard<std::mutex> lock(g->mu);
  return gb::shard_partition(g, parts, ranges);
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
  if (peer_count) {
    if (rg->num_chunks) gb::k_pr_pull<true><<<rg->grid_pull, gb::PR_THREADS, p->smem_bytes, s>>>(a);
    gb::k_pr_fix<true><<<rg->grid_fix, gb::PR_FIX_THREADS, 0, s>>>(a);
  } else {
    if (rg->num_chunks) gb::k_pr_pull<false><<<rg->grid_pull, gb::PR_THREADS, p->smem_bytes, s>>>(a);
    gb::k_pr_fix<false><<<rg->grid_fix, gb::PR_FIX_THREADS, 0, s>>>(a);
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
    // items = edges + rows
    uint64_t items = 0;
    if (rg->num_chunks) {
      uint2 c[2];
      gb::DeviceGuard guard(shard->graph->device);
      GB_CUDA(cudaMemcpy(&c[0], rg->coord.p, sizeof(uint2), cudaMemcpyDeviceToHost));
      GB_CUDA(cudaMemcpy(&c[1], rg->coord.p + rg->num_chunks, sizeof(uint2), cudaMemcpyDeviceToHost));
      items = (uint64_t)c[1].y - c[0].y;
    }
    *edges = items;
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

gb_status gb_page_rank_device(const gb_graph* graph, const gb_page_rank_config* config, float* d_scores,
                              uint64_t* ran_iterations, double* error) {
  GB_REQUIRE(d_scores != nullptr, "d_scores is NULL");
  return gb::page_rank_impl(graph, config, d_scores, nullptr, ran_iterations, error);
}

}  // extern "C"
