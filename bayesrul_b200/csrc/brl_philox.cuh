// Counter-based Philox4x32-10 noise, shared by every kernel.
// Keying (mirrors oracle/bnn_oracle.py::_philox_block, test infrastructure):
//   key     = (seed lo32, seed hi32)
//   counter = (element >> 2, global window index, global MC sample index, kind << 24 | layer/site)
// and element e takes lane (e & 3) of its block.  Normals are Box-Muller on lanes (0,1) and (2,3).
// Dropout keep-decisions use 16 bits each (8 per block) and a position-major element order
//   e = position * roundup8(C) + channel,  block = e >> 3,  half (e & 1) of word (e & 7) >> 1,
// so the 8 / 16 consecutive channels one epilogue thread owns at its position share 1 / 2 Philox blocks.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace brl {

enum NoiseKind : uint32_t {
  KIND_WEIGHT_EPS = 1,
  KIND_RADIAL_R = 2,
  KIND_LRT_EPS = 3,
  KIND_FLIP_IN = 4,
  KIND_FLIP_OUT = 5,
  KIND_DROPOUT = 6,
};

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}

__device__ __forceinline__ uint4 philox_block(uint64_t seed, uint32_t kind, uint32_t site, uint32_t sample,
                                              uint32_t window, uint32_t block) {
  return philox4x32_10(make_uint4(block, window, sample, (kind << 24) | (site & 0xFFFFFFu)),
                       make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
}

__device__ __forceinline__ float u01(uint32_t r) {  // (0,1), exact in fp32
  return ((float)(r >> 9) + 0.5f) * 1.1920928955078125e-07f;
}

__device__ __forceinline__ uint32_t lane_of(const uint4& r, uint32_t lane) {
  return lane == 0 ? r.x : lane == 1 ? r.y : lane == 2 ? r.z : r.w;
}

// four normals of one block
__device__ __forceinline__ float4 normal4(const uint4& r) {
  const float rad0 = sqrtf(-2.0f * logf(u01(r.x)));
  const float rad1 = sqrtf(-2.0f * logf(u01(r.z)));
  float s0, c0, s1, c1;
  sincosf(6.283185307179586f * u01(r.y), &s0, &c0);
  sincosf(6.283185307179586f * u01(r.w), &s1, &c1);
  return make_float4(rad0 * c0, rad0 * s0, rad1 * c1, rad1 * s1);
}

__device__ __forceinline__ float philox_normal(uint64_t seed, uint32_t kind, uint32_t site, uint32_t sample,
                                               uint32_t window, uint32_t elem) {
  const uint4 r = philox_block(seed, kind, site, sample, window, elem >> 2);
  const uint32_t lane = elem & 3u;
  const uint32_t a = lane < 2 ? r.x : r.z, b = lane < 2 ? r.y : r.w;
  const float rad = sqrtf(-2.0f * logf(u01(a)));
  float s, c;
  sincosf(6.283185307179586f * u01(b), &s, &c);
  return rad * ((lane & 1u) ? s : c);
}

__device__ __forceinline__ float philox_uniform(uint64_t seed, uint32_t kind, uint32_t site, uint32_t sample,
                                                uint32_t window, uint32_t elem) {
  const uint4 r = philox_block(seed, kind, site, sample, window, elem >> 2);
  return u01(lane_of(r, elem & 3u));
}

// 16-bit keep decision: ((h + 0.5) * 2^-16 < keep) on half `odd` of a Philox word
__device__ __forceinline__ bool keep16(uint32_t word, uint32_t odd, float keep) {
  const uint32_t h = odd ? (word >> 16) : (word & 0xFFFFu);
  return ((float)h + 0.5f) * 1.52587890625e-05f < keep;
}
__device__ __forceinline__ bool philox_keep(uint64_t seed, uint32_t kind, uint32_t site, uint32_t sample, uint32_t window,
                                            uint32_t C, uint32_t pos, uint32_t ch, float keep) {
  const uint32_t e = pos * ((C + 7u) & ~7u) + ch;
  const uint4 r = philox_block(seed, kind, site, sample, window, e >> 3);
  return keep16(lane_of(r, (e & 7u) >> 1), e & 1u, keep);
}

__device__ __forceinline__ float philox_sign(uint64_t seed, uint32_t kind, uint32_t site, uint32_t sample,
                                             uint32_t window, uint32_t elem) {
  const uint4 r = philox_block(seed, kind, site, sample, window, elem >> 2);
  return (lane_of(r, elem & 3u) >> 31) ? -1.0f : 1.0f;
}

}  // namespace brl
