/* cb_analysis.c — plan-time statistics of the column-blocked PageRank layout on the bench's R-MAT stream.
 *
 * Build:  gcc -O2 -fopenmp -o /tmp/cb_analysis tools/cb_analysis.c
 * Run:    /tmp/cb_analysis <scale> [block_entries=49152]
 *
 * Generates the same edge stream as orc_rmat_edges / rmat.cuh (the id scramble is a bijection and is
 * skipped: the PageRank plan renumbers by degree anyway), renumbers vertices like pagerank.cu
 * (in-degree desc, out-degree desc) and prints, for several hub thresholds T:
 *   - rows / edges with in-degree > T
 *   - cumulative share of those edges whose SOURCE lies in the first k source blocks
 *   - padded entry counts of the (32-row slice x block) layout with 8-entry groups
 * Design input for DESIGN.md §4.1; not part of the product or the tests. */
#include <omp.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static inline uint64_t rmat_mix(uint64_t z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

static void gen(uint32_t scale, uint64_t seed, uint64_t first, uint64_t count, uint32_t* src, uint32_t* dst) {
  const uint32_t A = 2448131358u, AB = 3264175144u, ABC = 4080218930u;
#pragma omp parallel for schedule(static)
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
    src[k] = s;
    dst[k] = t;
  }
}

static int cmp_u64(const void* a, const void* b) {
  uint64_t x = *(const uint64_t*)a, y = *(const uint64_t*)b;
  return x < y ? -1 : x > y;
}

int main(int argc, char** argv) {
  uint32_t scale = argc > 1 ? (uint32_t)atoi(argv[1]) : 22;
  uint32_t B = argc > 2 ? (uint32_t)atoi(argv[2]) : 49152;
  const uint64_t n = 1ull << scale, m = n * 16;
  uint32_t* src = malloc(m * 4);
  uint32_t* dst = malloc(m * 4);
  uint32_t* indeg = calloc(n, 4);
  uint32_t* outdeg = calloc(n, 4);
  double t0 = omp_get_wtime();
  gen(scale, 42, 0, m, src, dst);
#pragma omp parallel for schedule(static)
  for (uint64_t k = 0; k < m; ++k) {
    __atomic_fetch_add(&indeg[dst[k]], 1, __ATOMIC_RELAXED);
    __atomic_fetch_add(&outdeg[src[k]], 1, __ATOMIC_RELAXED);
  }
  fprintf(stderr, "generated %.1fs\n", omp_get_wtime() - t0);
  /* renumber: key = (~indeg, ~outdeg, id) ascending.  pack (~indeg:24? no) -> sort 64-bit (~indeg<<32|~outdeg)
   * plus id in a parallel array via sorting pairs of (key, id) as struct */
  typedef struct { uint64_t key; uint32_t id; } kv;
  /* qsort on 16-byte structs; key compare then id */
  uint64_t* keys = malloc(n * 16);
  for (uint64_t v = 0; v < n; ++v) {
    keys[2 * v] = ((uint64_t)(uint32_t)(~indeg[v]) << 32) | (uint32_t)(~outdeg[v]);
    keys[2 * v + 1] = v;
  }
  /* compare (key,id) lexicographically: both u64, reuse memcmp-free comparator */
  int cmp2(const void* a, const void* b) {
    const uint64_t* x = a; const uint64_t* y = b;
    if (x[0] != y[0]) return x[0] < y[0] ? -1 : 1;
    return x[1] < y[1] ? -1 : x[1] > y[1];
  }
  qsort(keys, n, 16, cmp2);
  uint32_t* new_id = malloc(n * 4);
  uint32_t* ideg = malloc(n * 4); /* in-degree by internal id */
  for (uint64_t r = 0; r < n; ++r) {
    new_id[keys[2 * r + 1]] = (uint32_t)r;
    ideg[r] = indeg[keys[2 * r + 1]];
  }
  free(keys);
  fprintf(stderr, "renumbered %.1fs\n", omp_get_wtime() - t0);
  /* relabel edges in place */
#pragma omp parallel for schedule(static)
  for (uint64_t k = 0; k < m; ++k) {
    src[k] = new_id[src[k]];
    dst[k] = new_id[dst[k]];
  }
  uint64_t n_active = 0;
  for (uint64_t r = 0; r < n; ++r) n_active += ideg[r] > 0;
  printf("scale %u n %lu m %lu n_active %lu block %u\n", scale, n, m, n_active, B);
  const uint32_t Ts[] = {16, 32, 64, 128, 256, 1024};
  const int NT = 6;
  uint64_t rows_gt[6] = {0}, edges_gt[6] = {0};
  for (uint64_t r = 0; r < n; ++r)
    for (int i = 0; i < NT; ++i)
      if (ideg[r] > Ts[i]) { rows_gt[i]++; edges_gt[i] += ideg[r]; }
  for (int i = 0; i < NT; ++i)
    printf("T>%u: rows %lu edges %lu (%.3f of m)\n", Ts[i], rows_gt[i], edges_gt[i], (double)edges_gt[i] / m);
  /* source-position distribution (all rows): share of edges with src < x */
  {
    const uint64_t xs[] = {8192, 16384, 32768, 49152, 65536, 131072, 262144, 524288, 1 << 20, 1 << 21, 1 << 22, 1 << 23, 1 << 24, 1 << 25};
    const int NX = 14;
    uint64_t c[14] = {0};
#pragma omp parallel
    {
      uint64_t lc[14] = {0};
#pragma omp for schedule(static)
      for (uint64_t k = 0; k < m; ++k)
        for (int i = 0; i < NX; ++i) lc[i] += src[k] < xs[i];
#pragma omp critical
      for (int i = 0; i < NX; ++i) c[i] += lc[i];
    }
    for (int i = 0; i < NX; ++i) printf("src<%lu: %.4f\n", xs[i], (double)c[i] / m);
  }
  /* per-threshold hub layout statistics */
  for (int ti = 1; ti < NT - 1; ++ti) {
    const uint32_t T = Ts[ti];
    const uint64_t n_long = rows_gt[ti];
    if (!n_long) continue;
    const uint64_t nblk_all = (n + B - 1) / B;
    /* count[(row, block)] for hub rows over the first KB blocks; rest = cold */
    const uint64_t KBs[] = {16, 32, 64, 128, 192, 256, 512};
    for (int ki = 0; ki < 7; ++ki) {
      uint64_t KB = KBs[ki];
      if (KB > nblk_all) KB = nblk_all;
      if ((double)n_long * KB * 2 > 24e9) continue;
      uint16_t* cnt = calloc(n_long * KB, 2); /* saturating not needed below 65535? hubs may exceed: use u32 if so */
      uint32_t* cnt32 = NULL;
      uint64_t hot_edges = 0, cold_edges = 0;
      /* need >16 bits for mega hubs in block 0: use u32 */
      free(cnt);
      cnt32 = calloc(n_long * KB, 4);
#pragma omp parallel for schedule(static) reduction(+ : hot_edges, cold_edges)
      for (uint64_t k = 0; k < m; ++k) {
        if (dst[k] >= n_long) continue;
        uint64_t b = src[k] / B;
        if (b < KB) {
          __atomic_fetch_add(&cnt32[(uint64_t)dst[k] * KB + b], 1, __ATOMIC_RELAXED);
          hot_edges++;
        } else cold_edges++;
      }
      /* padded entries per (row, block) segment: groups of 8 or 4 ids; LONG = segments >= 256 entries */
      uint64_t padded8 = 0, padded_row8 = 0, nonempty = 0, long_entries = 0;
#pragma omp parallel for schedule(static) reduction(+ : padded8, padded_row8, nonempty, long_entries)
      for (uint64_t r = 0; r < n_long; ++r)
        for (uint64_t b = 0; b < KB; ++b) {
          uint32_t c = cnt32[r * KB + b];
          padded_row8 += (c + 7) / 8 * 8;
          padded8 += (c + 3) / 4 * 4;
          nonempty += c > 0;
          if (c >= 256) long_entries += c;
        }
      printf("T>%u KB=%lu (src<%lu): hot_edges %lu (%.3f of hub, %.3f of m) cold %lu | pad4 x%.2f pad8 x%.2f pairs %lu (%.2f edges each) long-share %.2f\n",
             T, KB, KB * B, hot_edges, (double)hot_edges / (hot_edges + cold_edges), (double)hot_edges / m, cold_edges,
             (double)padded8 / hot_edges, (double)padded_row8 / hot_edges, nonempty,
             (double)hot_edges / nonempty, (double)long_entries / hot_edges);
      free(cnt32);
      if (KB == nblk_all) break;
    }
  }
  return 0;
}
