/*
 * oracle.c — CPU restatement of the neo4j-labs/graph hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library; the product (graph_b200/) never does.
 *
 * The reference is pure Rust and cannot be compiled in this image (no cargo/rustc), so there is
 * no oracle/_ref build.  Parity is pinned instead against every golden vector the reference's
 * own tests hold for this path (tests/golden/reference_goldens.json, tests/test_oracle_goldens.py).
 *
 * Each function cites the reference file:line it follows (paths relative to the reference
 * checkout).  Compile with -ffp-contract=off: rustc never contracts `a + b * c` into an FMA and
 * the PageRank goldens depend on that.
 */
#define _GNU_SOURCE
#include <float.h>
#include <math.h>
#include <pthread.h>
#include <stdatomic.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------ */
/* small helpers                                                                              */
/* ------------------------------------------------------------------------------------------ */

static int orc_threads(int requested) {
  if (requested > 0) return requested;
  long n = sysconf(_SC_NPROCESSORS_ONLN);
  return n > 0 ? (int)n : 4; /* DEFAULT_PARALLELISM = 4, crates/algos/src/lib.rs:152 */
}

ORC_API int orc_hardware_threads(void) { return orc_threads(0); }

static int cmp_u32(const void* a, const void* b) {
  uint32_t x = *(const uint32_t*)a, y = *(const uint32_t*)b;
  return (x > y) - (x < y);
}

typedef struct {
  uint32_t t;
  float w;
} orc_target; /* #[repr(C)] Target<u32,f32>, crates/builder/src/graph/mod.rs:6-10 */

static int cmp_target(const void* a, const void* b) {
  /* Target orders by `target` only (graph/mod.rs:12-30) */
  uint32_t x = ((const orc_target*)a)->t, y = ((const orc_target*)b)->t;
  return (x > y) - (x < y);
}

/* ------------------------------------------------------------------------------------------ */
/* input formats                                                                              */
/* ------------------------------------------------------------------------------------------ */

/* Graph500 packed edges: crates/builder/src/input/graph500.rs:111-127.
 * record = {v0_low, v1_low, high} little-endian u32; src = v0_low | (high & 0xFFFF) << 32,
 * dst = v1_low | (high >> 16) << 32.  Returns the edge count, or -1 if an id exceeds u32
 * (the reference panics in Idx::new, index.rs:51-54).  node_count = edge_count / 16 (:74). */
ORC_API int64_t orc_graph500_decode(const uint8_t* bytes, uint64_t len, uint32_t* src,
                                    uint32_t* dst) {
  uint64_t m = len / 12;
  for (uint64_t i = 0; i < m; ++i) {
    uint32_t rec[3];
    memcpy(rec, bytes + 12 * i, 12);
    uint64_t s = (uint64_t)rec[0] | ((uint64_t)(rec[2] & 0xFFFFu) << 32);
    uint64_t t = (uint64_t)rec[1] | ((uint64_t)(rec[2] >> 16) << 32);
    if (s > 0xFFFFFFFFull || t > 0xFFFFFFFFull) return -1;
    src[i] = (uint32_t)s;
    dst[i] = (uint32_t)t;
  }
  return (int64_t)m;
}

/* Text edge list "<src><sep><dst>[ <value>]\n" or "\r\n": edgelist.rs:181-279.
 * Pass 1 (src == NULL) counts edges; pass 2 fills.  ids base-10 (atoi), value f32, default 0. */
ORC_API int64_t orc_edgelist_parse(const char* text, uint64_t len, uint32_t* src, uint32_t* dst,
                                   float* val) {
  /* new_line_bytes, edgelist.rs:271-279 */
  uint64_t nl = 1;
  for (uint64_t i = 0; i < len; ++i)
    if (text[i] == '\n') {
      if (i > 0 && text[i - 1] == '\r') nl = 2;
      break;
    }
  uint64_t p = 0;
  int64_t m = 0;
  while (p < len) {
    uint64_t s = 0, t = 0;
    while (p < len && text[p] >= '0' && text[p] <= '9') s = s * 10 + (uint64_t)(text[p++] - '0');
    p += 1; /* one separator byte, edgelist.rs:225 */
    while (p < len && text[p] >= '0' && text[p] <= '9') t = t * 10 + (uint64_t)(text[p++] - '0');
    float v = 0.0f;
    if (p < len && text[p] == ' ') { /* optional value, edgelist.rs:230-236 */
      ++p;
      char* end = NULL;
      v = strtof(text + p, &end);
      p = (uint64_t)(end - text);
    }
    p += nl;
    if (src) {
      src[m] = (uint32_t)s;
      dst[m] = (uint32_t)t;
      if (val) val[m] = v;
    }
    ++m;
  }
  return m;
}

/* ------------------------------------------------------------------------------------------ */
/* CSR construction: crates/builder/src/graph/csr.rs:124-221                                  */
/* ------------------------------------------------------------------------------------------ */

enum { ORC_OUTGOING = 0, ORC_INCOMING = 1, ORC_UNDIRECTED = 2 };
enum { ORC_UNSORTED = 0, ORC_SORTED = 1, ORC_DEDUPLICATED = 2 };

/* Edges::max_node_id + 1, edgelist.rs:84-90 */
ORC_API uint32_t orc_node_count(const uint32_t* src, const uint32_t* dst, uint64_t m) {
  uint32_t mx = 0;
  for (uint64_t i = 0; i < m; ++i) {
    if (src[i] > mx) mx = src[i];
    if (dst[i] > mx) mx = dst[i];
  }
  return m ? mx + 1 : 0;
}

/* Builds offsets[n+1] (caller allocated) and returns the entry count.  targets/weights must
 * hold m entries (2m for undirected); weights may be NULL.  Single-thread order: edge-list
 * order, outgoing pass before incoming pass (csr.rs:154-172; the reference's own deterministic
 * test runs it on one thread, csr.rs:1195-1219).  Layouts: csr.rs:886-948. */
ORC_API uint64_t orc_csr_build(const uint32_t* src, const uint32_t* dst, const float* w,
                               uint64_t m, uint32_t n, int direction, int layout,
                               uint32_t* offsets, uint32_t* targets, float* weights) {
  uint64_t* deg = (uint64_t*)calloc((size_t)n + 1, sizeof(uint64_t));
  /* Edges::degrees, edgelist.rs:61-78 */
  for (uint64_t i = 0; i < m; ++i) {
    if (direction == ORC_OUTGOING || direction == ORC_UNDIRECTED) deg[src[i]]++;
    if (direction == ORC_INCOMING || direction == ORC_UNDIRECTED) deg[dst[i]]++;
  }
  /* prefix_sum_atomic, csr.rs:854-868 */
  uint64_t total = 0;
  for (uint32_t v = 0; v < n; ++v) {
    offsets[v] = (uint32_t)total;
    total += deg[v];
  }
  offsets[n] = (uint32_t)total;
  uint64_t* cur = deg; /* reuse as write cursors */
  for (uint32_t v = 0; v < n; ++v) cur[v] = offsets[v];
  orc_target* tmp = (orc_target*)malloc((size_t)(total ? total : 1) * sizeof(orc_target));
  if (direction == ORC_OUTGOING || direction == ORC_UNDIRECTED)
    for (uint64_t i = 0; i < m; ++i) {
      orc_target x = {dst[i], w ? w[i] : 0.0f};
      tmp[cur[src[i]]++] = x;
    }
  if (direction == ORC_INCOMING || direction == ORC_UNDIRECTED)
    for (uint64_t i = 0; i < m; ++i) {
      orc_target x = {src[i], w ? w[i] : 0.0f};
      tmp[cur[dst[i]]++] = x;
    }
  uint64_t out_len = total;
  if (layout == ORC_SORTED) {
    /* sort_targets, csr.rs:886-895.  qsort is not stable; equal targets with different values
     * have unspecified order in the reference too (sort_unstable). */
    for (uint32_t v = 0; v < n; ++v)
      qsort(tmp + offsets[v], offsets[v + 1] - offsets[v], sizeof(orc_target), cmp_target);
  } else if (layout == ORC_DEDUPLICATED) {
    /* sort_and_deduplicate_targets, csr.rs:897-948: sort, dedup, drop the row's own id */
    uint64_t wpos = 0;
    for (uint32_t v = 0; v < n; ++v) {
      uint64_t b = offsets[v], e = offsets[v + 1];
      qsort(tmp + b, e - b, sizeof(orc_target), cmp_target);
      offsets[v] = (uint32_t)wpos;
      for (uint64_t i = b; i < e; ++i) {
        if (i > b && tmp[i].t == tmp[i - 1].t) continue;
        if (tmp[i].t == v) continue;
        tmp[wpos++] = tmp[i];
      }
    }
    offsets[n] = (uint32_t)wpos;
    out_len = wpos;
  }
  for (uint64_t i = 0; i < out_len; ++i) {
    targets[i] = tmp[i].t;
    if (weights) weights[i] = tmp[i].w;
  }
  free(tmp);
  free(deg);
  return out_len;
}

/* make_degree_ordered: crates/builder/src/graph_ops.rs:511-638.
 * pairs (degree, id) sorted DESCENDING (:555 `left.cmp(right).reverse()`), so equal degrees put
 * the larger old id first; new_id[old] = rank (:564-592); rows rewritten with new ids and sorted
 * ascending (:595-638). new_id_out (n entries) may be NULL. */
typedef struct {
  uint32_t deg, id;
} orc_degpair;
static int cmp_degpair_desc(const void* a, const void* b) {
  const orc_degpair *x = (const orc_degpair*)a, *y = (const orc_degpair*)b;
  if (x->deg != y->deg) return (x->deg < y->deg) - (x->deg > y->deg);
  return (x->id < y->id) - (x->id > y->id);
}
ORC_API void orc_make_degree_ordered(const uint32_t* offsets, const uint32_t* targets, uint32_t n,
                                     uint32_t* new_offsets, uint32_t* new_targets,
                                     uint32_t* new_id_out) {
  orc_degpair* pairs = (orc_degpair*)malloc((size_t)(n ? n : 1) * sizeof(orc_degpair));
  uint32_t* new_id = (uint32_t*)malloc((size_t)(n ? n : 1) * sizeof(uint32_t));
  for (uint32_t v = 0; v < n; ++v) {
    pairs[v].deg = offsets[v + 1] - offsets[v];
    pairs[v].id = v;
  }
  qsort(pairs, n, sizeof(orc_degpair), cmp_degpair_desc);
  uint64_t total = 0;
  for (uint32_t r = 0; r < n; ++r) {
    new_id[pairs[r].id] = r;
    new_offsets[r] = (uint32_t)total;
    total += pairs[r].deg;
  }
  new_offsets[n] = (uint32_t)total;
  for (uint32_t u = 0; u < n; ++u) {
    uint32_t nu = new_id[u];
    uint64_t pos = new_offsets[nu];
    for (uint64_t i = offsets[u]; i < offsets[u + 1]; ++i) new_targets[pos++] = new_id[targets[i]];
    qsort(new_targets + new_offsets[nu], new_offsets[nu + 1] - new_offsets[nu], sizeof(uint32_t),
          cmp_u32);
  }
  if (new_id_out) memcpy(new_id_out, new_id, (size_t)n * sizeof(uint32_t));
  free(pairs);
  free(new_id);
}

/* in_degree_partition / greedy_node_map_partition: graph_ops.rs:431-439, :479-509.
 * Writes range boundaries into ranges[0..parts] and returns the number of ranges (<= parts). */
ORC_API uint32_t orc_in_degree_partition(const uint32_t* in_offsets, uint32_t n, uint64_t m,
                                         uint32_t parts, uint32_t* ranges) {
  uint64_t batch = (uint64_t)ceil((double)m / (double)parts);
  uint32_t count = 0;
  uint64_t acc = 0;
  uint32_t start = 0;
  ranges[0] = 0;
  for (uint32_t v = 0; v < n; ++v) {
    acc += in_offsets[v + 1] - in_offsets[v];
    if ((count < parts - 1 && acc >= batch) || v == n - 1) {
      ranges[++count] = v + 1;
      acc = 0;
      start = v + 1;
    }
  }
  (void)start;
  return count;
}

/* ------------------------------------------------------------------------------------------ */
/* PageRank: crates/algos/src/page_rank.rs:58-168                                             */
/* ------------------------------------------------------------------------------------------ */

/* The sweep exactly as ONE thread executes it (page_rank.rs:142-160): in place, sequential f32
 * sums in CSR order, separate multiply and add, IEEE division.  This is what the reference
 * computes whenever n <= CHUNK_SIZE (one chunk, one thread) — and it reproduces the goldens. */
ORC_API void orc_page_rank_seq(const uint32_t* in_off, const uint32_t* in_tgt,
                               const uint32_t* out_off, uint32_t n, uint64_t max_iterations,
                               double tolerance, float damping, float* scores,
                               uint64_t* ran_iterations, double* error_out) {
  float nf = (float)n;
  float init = 1.0f / nf;              /* page_rank.rs:70 */
  float base = (1.0f - damping) / nf;  /* page_rank.rs:71 */
  float* out = (float*)malloc((size_t)(n ? n : 1) * sizeof(float));
  for (uint32_t v = 0; v < n; ++v) {
    out[v] = init / (float)(out_off[v + 1] - out_off[v]); /* +inf for dangling, never read */
    scores[v] = init;
  }
  uint64_t it = 0;
  double err = 0.0;
  for (;;) {
    err = 0.0;
    for (uint32_t u = 0; u < n; ++u) {
      float tot = 0.0f;
      for (uint64_t e = in_off[u]; e < in_off[u + 1]; ++e) tot = tot + out[in_tgt[e]];
      float old = scores[u];
      float prod = damping * tot;
      float nw = base + prod;
      scores[u] = nw;
      float diff = nw - old;
      err += fabs((double)diff);
      out[u] = nw / (float)(out_off[u + 1] - out_off[u]);
    }
    ++it;
    if (err < tolerance || it == max_iterations) break; /* page_rank.rs:107 */
  }
  *ran_iterations = it;
  *error_out = err;
  free(out);
}

/* Same update rule on a double-buffered (Jacobi) schedule: every gather of sweep k reads the
 * out_scores written by sweep k-1.  acc64 != 0 accumulates the row sum in f64 (rounded once to
 * f32), which is the order-independent value the device kernel is gated against; acc64 == 0 keeps
 * the reference's sequential f32 adds. */
ORC_API void orc_page_rank_jacobi(const uint32_t* in_off, const uint32_t* in_tgt,
                                  const uint32_t* out_off, uint32_t n, uint64_t max_iterations,
                                  double tolerance, float damping, int acc64, float* scores,
                                  uint64_t* ran_iterations, double* error_out) {
  float nf = (float)n;
  float init = 1.0f / nf;
  float base = (1.0f - damping) / nf;
  float* cur = (float*)malloc((size_t)(n ? n : 1) * sizeof(float));
  float* nxt = (float*)malloc((size_t)(n ? n : 1) * sizeof(float));
  for (uint32_t v = 0; v < n; ++v) {
    cur[v] = init / (float)(out_off[v + 1] - out_off[v]);
    scores[v] = init;
  }
  uint64_t it = 0;
  double err = 0.0;
  for (;;) {
    err = 0.0;
    for (uint32_t u = 0; u < n; ++u) {
      float tot;
      if (acc64) {
        double t = 0.0;
        for (uint64_t e = in_off[u]; e < in_off[u + 1]; ++e) t += (double)cur[in_tgt[e]];
        tot = (float)t;
      } else {
        tot = 0.0f;
        for (uint64_t e = in_off[u]; e < in_off[u + 1]; ++e) tot = tot + cur[in_tgt[e]];
      }
      float old = scores[u];
      float prod = damping * tot;
      float nw = base + prod;
      scores[u] = nw;
      float diff = nw - old;
      err += fabs((double)diff);
      nxt[u] = nw / (float)(out_off[u + 1] - out_off[u]);
    }
    float* t = cur;
    cur = nxt;
    nxt = t;
    ++it;
    if (err < tolerance || it == max_iterations) break;
  }
  *ran_iterations = it;
  *error_out = err;
  free(cur);
  free(nxt);
}

/* The reference's multi-threaded schedule (page_rank.rs:113-168): T threads claim chunks of
 * CHUNK_SIZE = 16384 vertices with an atomic counter and update scores/out_scores IN PLACE
 * through raw shared pointers (an intentional race).  This is the TIMED CPU baseline. */
typedef struct {
  const uint32_t *in_off, *in_tgt, *out_off;
  uint32_t n;
  float base, damping;
  float *scores, *out;
  atomic_uint_least64_t* next_chunk;
  double err;
} pr_mt_arg;

static void* pr_mt_worker(void* p) {
  pr_mt_arg* a = (pr_mt_arg*)p;
  double err = 0.0;
  for (;;) {
    uint64_t start = atomic_fetch_add(a->next_chunk, 16384); /* page_rank.rs:12,135 */
    if (start >= a->n) break;
    uint64_t end = start + 16384 < a->n ? start + 16384 : a->n;
    for (uint64_t u = start; u < end; ++u) {
      float tot = 0.0f;
      for (uint64_t e = a->in_off[u]; e < a->in_off[u + 1]; ++e)
        tot = tot + ((volatile float*)a->out)[a->in_tgt[e]];
      float old = a->scores[u];
      float prod = a->damping * tot;
      float nw = a->base + prod;
      a->scores[u] = nw;
      float diff = nw - old;
      err += fabs((double)diff);
      ((volatile float*)a->out)[u] = nw / (float)(a->out_off[u + 1] - a->out_off[u]);
    }
  }
  a->err = err;
  return NULL;
}

typedef struct {
  const uint32_t* out_off;
  float *out, *scores;
  float init;
  uint64_t begin, end;
} pr_init_arg;
static void* pr_init_worker(void* p) {
  pr_init_arg* a = (pr_init_arg*)p;
  for (uint64_t v = a->begin; v < a->end; ++v) {
    a->out[v] = a->init / (float)(a->out_off[v + 1] - a->out_off[v]);
    a->scores[v] = a->init;
  }
  return NULL;
}

ORC_API void orc_page_rank_mt(const uint32_t* in_off, const uint32_t* in_tgt,
                              const uint32_t* out_off, uint32_t n, uint64_t max_iterations,
                              double tolerance, float damping, int threads, float* scores,
                              uint64_t* ran_iterations, double* error_out) {
  int T = orc_threads(threads);
  float nf = (float)n;
  float init = 1.0f / nf;
  float base = (1.0f - damping) / nf;
  float* out = (float*)malloc((size_t)(n ? n : 1) * sizeof(float));
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)T);
  pr_mt_arg* args = (pr_mt_arg*)malloc(sizeof(pr_mt_arg) * (size_t)T);
  { /* out_scores / scores initialised in parallel (par_iter, page_rank.rs:75-79) */
    pr_init_arg* ia = (pr_init_arg*)malloc(sizeof(pr_init_arg) * (size_t)T);
    for (int t = 0; t < T; ++t) {
      pr_init_arg a = {out_off, out, scores, init, (uint64_t)n * (uint64_t)t / (uint64_t)T,
                       (uint64_t)n * (uint64_t)(t + 1) / (uint64_t)T};
      ia[t] = a;
      pthread_create(&th[t], NULL, pr_init_worker, &ia[t]);
    }
    for (int t = 0; t < T; ++t) pthread_join(th[t], NULL);
    free(ia);
  }
  uint64_t it = 0;
  double err = 0.0;
  for (;;) {
    atomic_uint_least64_t next = 0;
    for (int t = 0; t < T; ++t) {
      pr_mt_arg a = {in_off, in_tgt, out_off, n, base, damping, scores, out, &next, 0.0};
      args[t] = a;
      pthread_create(&th[t], NULL, pr_mt_worker, &args[t]);
    }
    err = 0.0;
    for (int t = 0; t < T; ++t) {
      pthread_join(th[t], NULL);
      err += args[t].err; /* AtomicF64::fetch_add, page_rank.rs:162 (sum order unspecified) */
    }
    ++it;
    if (err < tolerance || it == max_iterations) break;
  }
  *ran_iterations = it;
  *error_out = err;
  free(th);
  free(args);
  free(out);
}

/* ------------------------------------------------------------------------------------------ */
/* WCC / Afforest: crates/algos/src/wcc.rs:158-301, afforest.rs:22-56                         */
/* ------------------------------------------------------------------------------------------ */

typedef _Atomic uint32_t au32;

/* Afforest::union, afforest.rs:22-39 */
static void af_union(au32* parent, uint32_t u, uint32_t v) {
  uint32_t p1 = atomic_load(&parent[u]);
  uint32_t p2 = atomic_load(&parent[v]);
  while (p1 != p2) {
    uint32_t high = p1 > p2 ? p1 : p2;
    uint32_t low = p1 + p2 - high;
    uint32_t p_high = atomic_load(&parent[high]);
    if (p_high == low) break;
    if (p_high == high) {
      uint32_t id = atomic_load(&parent[high]); /* self.find(high), afforest.rs:32 */
      uint32_t expect = high;
      if (atomic_compare_exchange_weak(&parent[id], &expect, low)) break;
    }
    p1 = atomic_load(&parent[atomic_load(&parent[high])]);
    p2 = atomic_load(&parent[low]);
  }
}

/* Afforest::compress for ids [b,e), afforest.rs:50-56 */
static void af_compress_range(au32* parent, uint32_t b, uint32_t e) {
  for (uint32_t x = b; x < e; ++x) {
    for (;;) {
      uint32_t p = atomic_load(&parent[x]);
      uint32_t pp = atomic_load(&parent[p]);
      if (p == pp) break;
      atomic_store(&parent[x], pp);
    }
  }
}

typedef struct {
  const uint32_t *out_off, *out_tgt, *in_off, *in_tgt;
  uint32_t n;
  au32* parent;
  uint64_t rounds;
  uint32_t skip;
  int phase; /* 0 sample, 1 compress, 2 link_remaining */
  uint64_t chunk;
  atomic_uint_least64_t* next;
} wcc_arg;

static void* wcc_worker(void* p) {
  wcc_arg* a = (wcc_arg*)p;
  for (;;) {
    uint64_t start = atomic_fetch_add(a->next, a->chunk);
    if (start >= a->n) break;
    uint64_t end = start + a->chunk < a->n ? start + a->chunk : a->n;
    if (a->phase == 1) {
      af_compress_range(a->parent, (uint32_t)start, (uint32_t)end);
      continue;
    }
    for (uint64_t u = start; u < end; ++u) {
      uint64_t ob = a->out_off[u], oe = a->out_off[u + 1];
      if (a->phase == 0) {
        /* sample_subgraph, wcc.rs:186-204: first `neighbor_rounds` out-edges */
        uint64_t lim = ob + a->rounds < oe ? ob + a->rounds : oe;
        for (uint64_t e = ob; e < lim; ++e) af_union(a->parent, (uint32_t)u, a->out_tgt[e]);
      } else {
        /* link_remaining, wcc.rs:274-301 */
        if (atomic_load(&a->parent[u]) == a->skip) continue;
        if (oe - ob > a->rounds)
          for (uint64_t e = ob + a->rounds; e < oe; ++e)
            af_union(a->parent, (uint32_t)u, a->out_tgt[e]);
        for (uint64_t e = a->in_off[u]; e < a->in_off[u + 1]; ++e)
          af_union(a->parent, (uint32_t)u, a->in_tgt[e]);
      }
    }
  }
  return NULL;
}

static void wcc_run_phase(wcc_arg* proto, int phase, int T) {
  atomic_uint_least64_t next = 0;
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)T);
  wcc_arg* args = (wcc_arg*)malloc(sizeof(wcc_arg) * (size_t)T);
  for (int t = 0; t < T; ++t) {
    args[t] = *proto;
    args[t].phase = phase;
    args[t].next = &next;
    if (T == 1) {
      wcc_worker(&args[t]);
    } else {
      pthread_create(&th[t], NULL, wcc_worker, &args[t]);
    }
  }
  if (T > 1)
    for (int t = 0; t < T; ++t) pthread_join(th[t], NULL);
  free(th);
  free(args);
}

static uint64_t splitmix64(uint64_t* s) {
  uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

/* wcc_afforest(graph, config).to_vec(): wcc.rs:127-139,158-183; afforest.rs:100-114.
 * threads == 1 runs every phase on the calling thread in id order.  The random sample
 * (wcc.rs:245-271; the reference seeds WyRand from entropy) only chooses which component is
 * skipped in link_remaining; the result never depends on it. */
ORC_API void orc_wcc_afforest(const uint32_t* out_off, const uint32_t* out_tgt,
                              const uint32_t* in_off, const uint32_t* in_tgt, uint32_t n,
                              uint64_t chunk_size, uint64_t neighbor_rounds, uint64_t sampling_size,
                              uint64_t rng_seed, int threads, uint32_t* components) {
  int T = orc_threads(threads);
  au32* parent = (au32*)malloc((size_t)(n ? n : 1) * sizeof(au32));
  for (uint32_t i = 0; i < n; ++i) atomic_init(&parent[i], i); /* Afforest::new, afforest.rs:76-86 */
  wcc_arg proto = {out_off, out_tgt, in_off, in_tgt, n, parent, neighbor_rounds, 0, 0,
                   chunk_size ? chunk_size : 16384, NULL};
  wcc_run_phase(&proto, 0, T); /* sample_subgraph */
  wcc_run_phase(&proto, 1, T); /* compress */
  /* find_largest_component, wcc.rs:245-271 */
  uint32_t best = 0;
  if (n > 0 && sampling_size > 0) {
    uint32_t* samp = (uint32_t*)malloc((size_t)sampling_size * sizeof(uint32_t));
    uint64_t s = rng_seed;
    for (uint64_t i = 0; i < sampling_size; ++i)
      samp[i] = atomic_load(&parent[splitmix64(&s) % n]);
    qsort(samp, sampling_size, sizeof(uint32_t), cmp_u32);
    uint64_t best_cnt = 0, run = 0;
    for (uint64_t i = 0; i < sampling_size; ++i) {
      run = (i > 0 && samp[i] == samp[i - 1]) ? run + 1 : 1;
      if (run > best_cnt) {
        best_cnt = run;
        best = samp[i];
      }
    }
    free(samp);
  }
  proto.skip = best;
  wcc_run_phase(&proto, 2, T); /* link_remaining */
  wcc_run_phase(&proto, 1, T); /* compress */
  for (uint32_t i = 0; i < n; ++i) components[i] = atomic_load(&parent[i]);
  free(parent);
}

/* ------------------------------------------------------------------------------------------ */
/* DisjointSetStruct: crates/algos/src/dss.rs:38-116 (single thread)                          */
/* ------------------------------------------------------------------------------------------ */
/* find with path halving, dss.rs:76-94: every visited id is pointed at its grand parent */
static uint32_t dss_find(uint32_t* p, uint32_t id) {
  uint32_t parent = p[id];
  while (id != parent) {
    uint32_t grand = p[parent];
    if (p[id] == parent) p[id] = grand; /* update_parent(id, parent, grand_parent): CAS, result ignored */
    id = parent;
    parent = grand;
  }
  return id;
}
/* union by min, dss.rs:38-62 */
static void dss_union(uint32_t* p, uint32_t id1, uint32_t id2) {
  for (;;) {
    id1 = dss_find(p, id1);
    id2 = dss_find(p, id2);
    if (id1 == id2) return;
    if (id1 < id2) {
      uint32_t t = id1;
      id1 = id2;
      id2 = t;
    }
    if (p[id1] == id1) { /* update_parent(id1, id1, id2) */
      p[id1] = id2;
      return;
    }
  }
}
ORC_API void orc_dss_ops(uint32_t n, const uint32_t* pairs, uint64_t npairs, uint32_t* parents, uint32_t* finds) {
  /* unions in order, then find(i) for every i (dss.rs tests :183-220); parents = to_vec() BEFORE the finds */
  for (uint32_t i = 0; i < n; ++i) parents[i] = i;
  for (uint64_t k = 0; k < npairs; ++k) dss_union(parents, pairs[2 * k], pairs[2 * k + 1]);
  uint32_t* work = (uint32_t*)malloc((size_t)(n ? n : 1) * 4);
  memcpy(work, parents, (size_t)n * 4);
  for (uint32_t i = 0; i < n; ++i) finds[i] = dss_find(work, i);
  free(work);
}
/* wcc_baseline (variant 0, wcc.rs:103-123: union over every out-edge, no compress) and
 * wcc_afforest_dss (variant 1, wcc.rs:144-156 + wcc() :158-183 with the DSS as union-find), one thread
 * in id order.  to_vec = the raw parent array the reference's to_vec() returns (dss.rs:158: entries may
 * be non-root ancestors: path halving does not fully compress); component[i] = find(i) = what
 * Components::component returns = the minimum node id of i's component. */
ORC_API void orc_wcc_dss(const uint32_t* out_off, const uint32_t* out_tgt, const uint32_t* in_off,
                         const uint32_t* in_tgt, uint32_t n, int variant, uint64_t neighbor_rounds,
                         uint64_t sampling_size, uint64_t rng_seed, uint32_t* to_vec, uint32_t* component) {
  uint32_t* p = to_vec;
  for (uint32_t i = 0; i < n; ++i) p[i] = i;
  if (variant == 0) {
    for (uint32_t u = 0; u < n; ++u)
      for (uint64_t e = out_off[u]; e < out_off[u + 1]; ++e) dss_union(p, u, out_tgt[e]);
  } else {
    for (uint32_t u = 0; u < n; ++u) { /* sample_subgraph, wcc.rs:186-204 */
      uint64_t b = out_off[u], e = out_off[u + 1];
      uint64_t lim = (e - b < neighbor_rounds) ? e : b + neighbor_rounds;
      for (uint64_t i = b; i < lim; ++i) dss_union(p, u, out_tgt[i]);
    }
    for (uint32_t i = 0; i < n; ++i) dss_find(p, i); /* compress, dss.rs:112-116 */
    uint32_t best = 0;                                /* find_largest_component, wcc.rs:245-271 */
    if (n > 0 && sampling_size > 0) {
      uint32_t* samp = (uint32_t*)malloc((size_t)sampling_size * sizeof(uint32_t));
      uint64_t s = rng_seed;
      for (uint64_t i = 0; i < sampling_size; ++i) samp[i] = dss_find(p, (uint32_t)(splitmix64(&s) % n));
      qsort(samp, sampling_size, sizeof(uint32_t), cmp_u32);
      uint64_t best_cnt = 0, run = 0;
      for (uint64_t i = 0; i < sampling_size; ++i) {
        run = (i > 0 && samp[i] == samp[i - 1]) ? run + 1 : 1;
        if (run > best_cnt) {
          best_cnt = run;
          best = samp[i];
        }
      }
      free(samp);
    }
    for (uint32_t u = 0; u < n; ++u) { /* link_remaining, wcc.rs:274-301 */
      if (sampling_size > 0 && dss_find(p, u) == best) continue;
      uint64_t b = out_off[u], e = out_off[u + 1];
      if (e - b > neighbor_rounds)
        for (uint64_t i = b + neighbor_rounds; i < e; ++i) dss_union(p, u, out_tgt[i]);
      for (uint64_t i = in_off[u]; i < in_off[u + 1]; ++i) dss_union(p, u, in_tgt[i]);
    }
    for (uint32_t i = 0; i < n; ++i) dss_find(p, i); /* final compress */
  }
  uint32_t* work = (uint32_t*)malloc((size_t)(n ? n : 1) * 4);
  memcpy(work, p, (size_t)n * 4);
  for (uint32_t i = 0; i < n; ++i) component[i] = dss_find(work, i);
  free(work);
}

/* Independent statement of the result: label = minimum node id of the weakly connected
 * component (the invariant parent[x] <= x of afforest.rs:22-39 plus full compression). */
ORC_API void orc_wcc_min_label(const uint32_t* out_off, const uint32_t* out_tgt, uint32_t n,
                               uint32_t* components) {
  uint32_t* p = components;
  for (uint32_t i = 0; i < n; ++i) p[i] = i;
  for (uint32_t u = 0; u < n; ++u)
    for (uint64_t e = out_off[u]; e < out_off[u + 1]; ++e) {
      uint32_t a = u, b = out_tgt[e];
      while (p[a] != a) a = p[a] = p[p[a]];
      while (p[b] != b) b = p[b] = p[p[b]];
      if (a < b) p[b] = a;
      else if (b < a) p[a] = b;
    }
  for (uint32_t i = 0; i < n; ++i) {
    uint32_t r = i;
    while (p[r] != r) r = p[r];
    p[i] = r;
  }
}

/* ------------------------------------------------------------------------------------------ */
/* SSSP / delta-stepping: crates/algos/src/sssp.rs:38-204                                     */
/* ------------------------------------------------------------------------------------------ */

typedef struct {
  uint32_t* v;
  uint64_t len, cap;
} orc_bin;
typedef struct {
  orc_bin* bins;
  uint64_t len;
} orc_bins; /* ThreadLocalBins, sssp.rs:228-275 */

static void bins_resize(orc_bins* b, uint64_t new_len) {
  if (new_len <= b->len) return;
  b->bins = (orc_bin*)realloc(b->bins, (size_t)new_len * sizeof(orc_bin));
  for (uint64_t i = b->len; i < new_len; ++i) {
    b->bins[i].v = NULL;
    b->bins[i].len = b->bins[i].cap = 0;
  }
  b->len = new_len;
}
static void bin_push(orc_bin* b, uint32_t x) {
  if (b->len == b->cap) {
    b->cap = b->cap ? 2 * b->cap : 16;
    b->v = (uint32_t*)realloc(b->v, (size_t)b->cap * sizeof(uint32_t));
  }
  b->v[b->len++] = x;
}

/* relax_edges, sssp.rs:170-204 (single thread: the CAS always succeeds) */
static void ds_relax(const uint32_t* off, const uint32_t* tgt, const float* w, float* dist,
                     orc_bins* bins, uint32_t node, float delta) {
  for (uint64_t e = off[node]; e < off[node + 1]; ++e) {
    uint32_t t = tgt[e];
    float old = dist[t];
    float nd = dist[node] + w[e];
    if (nd < old) {
      dist[t] = nd;
      float q = nd / delta;
      /* `as usize` saturates, sssp.rs:190 */
      uint64_t dest = q >= 1.8446744e19f ? UINT64_MAX : (q > 0.0f ? (uint64_t)q : 0);
      if (dest >= bins->len) bins_resize(bins, dest + 1);
      bin_push(&bins->bins[dest], t);
    }
  }
}

/* delta_stepping with one rayon thread (sssp.rs:38-102): shared frontier, local bins, the
 * BIN_SIZE_THRESHOLD = 1000 local drain (sssp.rs:134-157).  Returns 0, or -1 on a bad start. */
ORC_API int orc_sssp_delta_stepping(const uint32_t* off, const uint32_t* tgt, const float* w,
                                    uint32_t n, uint64_t start, float delta, float* dist) {
  if (start >= n) return -1;
  for (uint32_t i = 0; i < n; ++i) dist[i] = FLT_MAX; /* INF = f32::MAX, sssp.rs:12 */
  dist[start] = 0.0f;
  uint64_t m = off[n];
  uint32_t* frontier = (uint32_t*)malloc((size_t)(m ? m : 1) * sizeof(uint32_t));
  frontier[0] = (uint32_t)start;
  uint64_t frontier_len = 1;
  orc_bins bins = {NULL, 0};
  bins_resize(&bins, 1);
  uint64_t curr = 0;
  const uint64_t NO_BIN = UINT64_MAX;
  while (curr != NO_BIN) {
    /* process_shared_bin, sssp.rs:104-132 */
    for (uint64_t i = 0; i < frontier_len; ++i) {
      uint32_t node = frontier[i];
      if (dist[node] >= delta * (float)curr) ds_relax(off, tgt, w, dist, &bins, node, delta);
    }
    /* process_local_bins, sssp.rs:134-157 */
    while (curr < bins.len && bins.bins[curr].len != 0 && bins.bins[curr].len < 1000) {
      orc_bin copy = bins.bins[curr];
      uint32_t* tmp = (uint32_t*)malloc((size_t)copy.len * sizeof(uint32_t));
      memcpy(tmp, copy.v, (size_t)copy.len * sizeof(uint32_t));
      uint64_t cl = copy.len;
      bins.bins[curr].len = 0;
      for (uint64_t i = 0; i < cl; ++i) ds_relax(off, tgt, w, dist, &bins, tmp[i], delta);
      free(tmp);
    }
    /* min_non_empty_bin, sssp.rs:159-168 */
    uint64_t next = NO_BIN;
    for (uint64_t b = curr; b < bins.len; ++b)
      if (bins.bins[b].len != 0) {
        next = b;
        break;
      }
    /* copy the next bin into the shared frontier, sssp.rs:82-94 */
    frontier_len = 0;
    if (next != NO_BIN) {
      orc_bin* nb = &bins.bins[next];
      if (nb->len > m) { /* the reference's frontier has edge_count slots (sssp.rs:54) */
        frontier = (uint32_t*)realloc(frontier, (size_t)nb->len * sizeof(uint32_t));
        m = nb->len;
      }
      memcpy(frontier, nb->v, (size_t)nb->len * sizeof(uint32_t));
      frontier_len = nb->len;
      nb->len = 0;
    }
    curr = next;
  }
  for (uint64_t b = 0; b < bins.len; ++b) free(bins.bins[b].v);
  free(bins.bins);
  free(frontier);
  return 0;
}

/* Independent statement of the fixed point: label-correcting Bellman-Ford sweeps in f32 until
 * nothing changes.  dist[t] = min_u fl32(dist[u] + w(u,t)) is unique for w >= 0 because f32 `+`
 * is monotone (SURVEY.md §7 hard part 7). */
ORC_API int orc_sssp_bellman_ford(const uint32_t* off, const uint32_t* tgt, const float* w,
                                  uint32_t n, uint64_t start, float* dist) {
  if (start >= n) return -1;
  for (uint32_t i = 0; i < n; ++i) dist[i] = FLT_MAX;
  dist[start] = 0.0f;
  int changed = 1;
  while (changed) {
    changed = 0;
    for (uint32_t u = 0; u < n; ++u) {
      if (dist[u] == FLT_MAX) continue;
      for (uint64_t e = off[u]; e < off[u + 1]; ++e) {
        float nd = dist[u] + w[e];
        if (nd < dist[tgt[e]]) {
          dist[tgt[e]] = nd;
          changed = 1;
        }
      }
    }
  }
  return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* triangle count: crates/algos/src/triangle_count.rs:22-86 (+ utils.rs:8-101 put-back)       */
/* ------------------------------------------------------------------------------------------ */

static uint64_t tc_vertex(const uint32_t* off, const uint32_t* tgt, uint32_t u) {
  uint64_t tri = 0;
  uint64_t ub = off[u], ue = off[u + 1];
  for (uint64_t i = ub; i < ue; ++i) {
    uint32_t v = tgt[i];
    if (v > u) break; /* triangle_count.rs:49-51 */
    uint64_t it = ub; /* fresh put-back iterator over N(u), :53 */
    for (uint64_t j = off[v]; j < off[v + 1]; ++j) {
      uint32_t w = tgt[j];
      if (w > v) break; /* :56-58 */
      /* advance while x < w; an x >= w is put back, i.e. the cursor stays on it (:59-66) */
      while (it < ue && tgt[it] < w) ++it;
      if (it == ue) {
        /* iterator exhausted: later w find nothing, but the reference keeps scanning N(v);
         * the count cannot change any more */
        break;
      }
      if (tgt[it] == w) ++tri;
    }
  }
  return tri;
}

typedef struct {
  const uint32_t *off, *tgt;
  uint32_t n;
  atomic_uint_least64_t* next;
  uint64_t tri;
} tc_arg;

static void* tc_worker(void* p) {
  tc_arg* a = (tc_arg*)p;
  uint64_t tri = 0;
  for (;;) {
    uint64_t start = atomic_fetch_add(a->next, 64); /* CHUNK_SIZE = 64, triangle_count.rs:10,40 */
    if (start >= a->n) break;
    uint64_t end = start + 64 < a->n ? start + 64 : a->n;
    for (uint64_t u = start; u < end; ++u) tri += tc_vertex(a->off, a->tgt, (uint32_t)u);
  }
  a->tri = tri;
  return NULL;
}

ORC_API uint64_t orc_triangle_count(const uint32_t* off, const uint32_t* tgt, uint32_t n,
                                    int threads) {
  int T = orc_threads(threads);
  atomic_uint_least64_t next = 0;
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)T);
  tc_arg* args = (tc_arg*)malloc(sizeof(tc_arg) * (size_t)T);
  for (int t = 0; t < T; ++t) {
    tc_arg a = {off, tgt, n, &next, 0};
    args[t] = a;
    if (T == 1) tc_worker(&args[t]);
    else pthread_create(&th[t], NULL, tc_worker, &args[t]);
  }
  uint64_t total = 0;
  for (int t = 0; t < T; ++t) {
    if (T > 1) pthread_join(th[t], NULL);
    total += args[t].tri;
  }
  free(th);
  free(args);
  return total;
}

/* ------------------------------------------------------------------------------------------ */
/* synthetic R-MAT stream (shared definition with graph_b200/csrc/rmat.cuh)                    */
/* ------------------------------------------------------------------------------------------ */
/* Not part of the reference (it reads Graph500 files, input/graph500.rs); this is the
 * workload generator BASELINE.json names.  Pure integer arithmetic so that CPU and GPU emit
 * identical edges: edge i draws `scale` quadrant choices from a counter-based splitmix64 stream
 * keyed by (seed, i) with thresholds a,b,c = .57,.19,.19 on 32-bit draws, then both endpoints go
 * through a fixed bijective scramble of the `scale`-bit id space. */
static inline uint64_t rmat_mix(uint64_t z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
static inline uint32_t rmat_scramble(uint32_t v, uint32_t scale, uint64_t seed) {
  uint32_t mask = scale >= 32 ? 0xFFFFFFFFu : ((1u << scale) - 1u);
  uint32_t k0 = (uint32_t)rmat_mix(seed ^ 0xA5A5A5A5DEADBEEFull) | 1u;
  uint32_t k1 = (uint32_t)(rmat_mix(seed ^ 0x0123456789ABCDEFull) >> 32);
  uint32_t h = scale / 2 ? scale / 2 : 1;
  v = (v * k0) & mask;  /* odd multiplier: bijection mod 2^scale */
  v ^= v >> h;          /* xorshift: bijection */
  v = (v + k1) & mask;
  v = (v * 0x9E3779B1u) & mask;
  v ^= v >> h;
  return v & mask;
}
ORC_API void orc_rmat_edges(uint32_t scale, uint64_t seed, uint64_t first, uint64_t count,
                            uint32_t* src, uint32_t* dst) {
  const uint32_t A = 2448131358u;   /* floor(0.57 * 2^32) */
  const uint32_t AB = 3264175144u;  /* floor(0.76 * 2^32) */
  const uint32_t ABC = 4080218930u; /* floor(0.95 * 2^32) */
  for (uint64_t k = 0; k < count; ++k) {
    uint64_t i = first + k;
    uint64_t state = rmat_mix(seed + 0x9E3779B97F4A7C15ull * (i + 1));
    uint32_t s = 0, t = 0;
    for (uint32_t level = 0; level < scale; level += 2) {
      state += 0x9E3779B97F4A7C15ull;
      uint64_t z = rmat_mix(state);
      uint32_t r0 = (uint32_t)(z >> 32), r1 = (uint32_t)z;
      s = (s << 1) | (uint32_t)(r0 >= AB);
      t = (t << 1) | (uint32_t)((r0 >= A && r0 < AB) || r0 >= ABC);
      if (level + 1 < scale) {
        s = (s << 1) | (uint32_t)(r1 >= AB);
        t = (t << 1) | (uint32_t)((r1 >= A && r1 < AB) || r1 >= ABC);
      }
    }
    src[k] = rmat_scramble(s, scale, seed);
    dst[k] = rmat_scramble(t, scale, seed);
  }
}
/* deterministic uniform (0,1] f32 weight of edge i (for SSSP workloads) */
ORC_API void orc_rmat_weights(uint64_t seed, uint64_t first, uint64_t count, float* w) {
  for (uint64_t k = 0; k < count; ++k) {
    uint64_t z = rmat_mix((seed ^ 0x5851F42D4C957F2Dull) + 0x9E3779B97F4A7C15ull * (first + k + 1));
    w[k] = (float)((uint32_t)(z >> 40) + 1u) * (1.0f / 16777216.0f); /* (0,1], 24-bit exact */
  }
}

/* wall-clock helper for the CPU baseline legs */
ORC_API double orc_now_seconds(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}
