// rmat.cuh — the synthetic R-MAT / Graph500-style edge stream BASELINE.json names.
// Bit-identical twin of oracle/oracle.c:orc_rmat_edges (pure integer arithmetic): edge i draws
// `scale` quadrant choices (a,b,c,d = .57,.19,.19,.05) from a counter-based splitmix64 stream
// keyed by (seed, i); both endpoints then pass a fixed bijective scramble of the scale-bit ids.
// The reference itself only READS such graphs (crates/builder/src/input/graph500.rs:63-127).
#pragma once
#include <cstdint>

namespace gb {

__host__ __device__ __forceinline__ uint64_t rmat_mix(uint64_t z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

struct RmatScramble {
  uint32_t mask, k0, k1, h;
  __host__ __device__ RmatScramble(uint32_t scale, uint64_t seed) {
    mask = scale >= 32 ? 0xFFFFFFFFu : ((1u << scale) - 1u);
    k0 = static_cast<uint32_t>(rmat_mix(seed ^ 0xA5A5A5A5DEADBEEFull)) | 1u;
    k1 = static_cast<uint32_t>(rmat_mix(seed ^ 0x0123456789ABCDEFull) >> 32);
    h = scale / 2 ? scale / 2 : 1;
  }
  __host__ __device__ __forceinline__ uint32_t operator()(uint32_t v) const {
    v = (v * k0) & mask;
    v ^= v >> h;
    v = (v + k1) & mask;
    v = (v * 0x9E3779B1u) & mask;
    v ^= v >> h;
    return v & mask;
  }
};

__host__ __device__ __forceinline__ void rmat_edge(uint32_t scale, uint64_t seed, uint64_t i,
                                                   const RmatScramble& scr, uint32_t* src,
                                                   uint32_t* dst) {
  const uint32_t A = 2448131358u;    // floor(0.57 * 2^32)
  const uint32_t AB = 3264175144u;   // floor(0.76 * 2^32)
  const uint32_t ABC = 4080218930u;  // floor(0.95 * 2^32)
  uint64_t state = rmat_mix(seed + 0x9E3779B97F4A7C15ull * (i + 1));
  uint32_t s = 0, t = 0;
  for (uint32_t level = 0; level < scale; level += 2) {
    state += 0x9E3779B97F4A7C15ull;
    uint64_t z = rmat_mix(state);
    uint32_t r0 = static_cast<uint32_t>(z >> 32), r1 = static_cast<uint32_t>(z);
    s = (s << 1) | static_cast<uint32_t>(r0 >= AB);
    t = (t << 1) | static_cast<uint32_t>((r0 >= A && r0 < AB) || r0 >= ABC);
    if (level + 1 < scale) {
      s = (s << 1) | static_cast<uint32_t>(r1 >= AB);
      t = (t << 1) | static_cast<uint32_t>((r1 >= A && r1 < AB) || r1 >= ABC);
    }
  }
  *src = scr(s);
  *dst = scr(t);
}

__host__ __device__ __forceinline__ float rmat_weight(uint64_t seed, uint64_t i) {
  uint64_t z = rmat_mix((seed ^ 0x5851F42D4C957F2Dull) + 0x9E3779B97F4A7C15ull * (i + 1));
  return static_cast<float>(static_cast<uint32_t>(z >> 40) + 1u) * (1.0f / 16777216.0f);
}

}  // namespace gb
