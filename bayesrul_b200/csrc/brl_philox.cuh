// Counter-based Philox4x32-10 noise, shared by every kernel.
// Keying (mirrors oracle/bnn_oracle.py::_philox_block, test infrastructure):
//   key     = (seed lo32, seed hi32)
//   counter = (element >> 2, global window index, global MC sample index, kind << 24 | layer/site)
// and element e takes lane (e & 3) of its block.  Normals are Box-Muller on lanes (0,1) and (2,3).
// Dropout keep-decisions: SIXTEEN per block, 14-bit resolution.  Element order is position-major,
//   e = position * roundup16(C) + channel,  block = e >> 4,  decision bi = (e & 12) | {0, 2, 1, 3}[e & 3] of the block
// (the middle two of every four are swapped so that the two 16-bit lanes of one SIMD compare are two ADJACENT channels,
// i.e. one packed half2 of the fused kernel is masked with a single AND), so the 8 / 16 consecutive channels one
// epilogue thread owns at its position share one Philox block.  With B[0..15] the
// bytes of the block (little-endian words x, y, z, w), decision bi compares the 14-bit number
//   u = (B[(bi + 1) & 15] << 8 | B[bi]) & 0x3FFF      with      T = min(ceil(keep * 2^14 - 0.5), 16383)   (keep iff u < T),
// i.e. the low 14 bits of the sixteen overlapping 16-bit windows of the 128 bits: every decision has exactly the 14-bit marginal
// T / 2^14 (|T / 2^14 - keep| <= 2^-14), and two neighbouring decisions share only bits that are the LOW byte of one of them (their
// covariance is < 2^-8 of a Bernoulli variance; tests/test_host_logic.py bounds the empirical correlations).  A 14-bit number is the bit
// pattern of a non-negative finite fp16 (exponent < 16) whose float order is its integer order, so two decisions are ONE packed-half
// compare with a mask result (HSET2) behind one AND: eight pairs per block, branch-free.  (Round 1 compared 16-bit lanes with the
// emulated integer SIMD compare -- six ALU-pipe instructions per pair, which made the MC-dropout epilogue ALU-pipe-bound.)
#pragma once
#include <cstdint>
#include <cuda_fp16.h>
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

template <int ROUNDS>
__device__ __forceinline__ uint4 philox4x32(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < ROUNDS; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) { return philox4x32<10>(c, k); }

__device__ __forceinline__ uint4 philox_block(uint64_t seed, uint32_t kind, uint32_t site, uint32_t sample,
                                              uint32_t window, uint32_t block) {
  return philox4x32<10>(make_uint4(block, window, sample, (kind << 24) | (site & 0xFFFFFFu)),
                        make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
}
// Dropout masks use Philox4x32-7: seven rounds are the smallest Crush-resistant member of the family (Salmon et al., SC'11;
// ten is its safety margin), one mask bit decision consumes 8 of the 128 output bits, and the generator sits on the critical
// path of the fused kernel's epilogue warps.  Normals and signs keep ten rounds.
__device__ __forceinline__ uint4 philox_block_mask(uint64_t seed, uint32_t kind, uint32_t site, uint32_t sample,
                                                   uint32_t window, uint32_t block) {
  return philox4x32<7>(make_uint4(block, window, sample, (kind << 24) | (site & 0xFFFFFFu)),
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

// integer threshold of a keep probability: u < T  <=>  (u + 0.5) * 2^-14 < keep   (u a 14-bit number)
__host__ __device__ __forceinline__ uint32_t keep_threshold(float keep) {
  const uint32_t t = (uint32_t)ceilf(keep * 16384.0f - 0.5f);
  return t > 16383u ? 16383u : t;
}
// all sixteen decisions of block r: ev[i] / od[i] hold decisions 4i, 4i+2 / 4i+1, 4i+3 as all-ones 16-bit lanes
struct KeepBits { uint32_t ev[4], od[4]; };
__device__ __forceinline__ uint32_t keep_lt2(uint32_t w, uint32_t T2) {  // both 16-bit lanes: (lane & 0x3FFF) < T ? 0xFFFF : 0
  const uint32_t h = w & 0x3FFF3FFFu;
  return __hlt2_mask(*reinterpret_cast<const __half2*>(&h), *reinterpret_cast<const __half2*>(&T2));
}
__device__ __forceinline__ KeepBits keep_bits_packed(const uint4& r, uint32_t T2) {  // T2 = T | T << 16
  KeepBits k;
  k.ev[0] = keep_lt2(r.x, T2); k.ev[1] = keep_lt2(r.y, T2); k.ev[2] = keep_lt2(r.z, T2); k.ev[3] = keep_lt2(r.w, T2);
  k.od[0] = keep_lt2(__funnelshift_r(r.x, r.y, 8), T2);
  k.od[1] = keep_lt2(__funnelshift_r(r.y, r.z, 8), T2);
  k.od[2] = keep_lt2(__funnelshift_r(r.z, r.w, 8), T2);
  k.od[3] = keep_lt2(__funnelshift_r(r.w, r.x, 8), T2);
  return k;
}
__device__ __forceinline__ KeepBits keep_bits(const uint4& r, uint32_t T) { return keep_bits_packed(r, T | (T << 16)); }
// all-ones / all-zeros 16-bit lanes for the channel pair (2p, 2p + 1) of the block's sixteen channels, p = 0..7: channels
// 4i, 4i+1 take decisions 4i, 4i+2 (the lanes of ev[i]); channels 4i+2, 4i+3 take decisions 4i+1, 4i+3 (the lanes of od[i])
__device__ __forceinline__ uint32_t keep_pair(const KeepBits& k, uint32_t p) {
  const uint32_t i = p >> 1;
  return (p & 1u) ? (i == 0 ? k.od[0] : i == 1 ? k.od[1] : i == 2 ? k.od[2] : k.od[3])
                  : (i == 0 ? k.ev[0] : i == 1 ? k.ev[1] : i == 2 ? k.ev[2] : k.ev[3]);
}
// channel c (0..15) of the block
__device__ __forceinline__ bool keep_at(const KeepBits& k, uint32_t c) { return (keep_pair(k, c >> 1) >> (16u * (c & 1u))) & 1u; }
__device__ __forceinline__ bool philox_keep(uint64_t seed, uint32_t kind, uint32_t site, uint32_t sample, uint32_t window,
                                            uint32_t C, uint32_t pos, uint32_t ch, float keep) {
  const uint32_t e = pos * ((C + 15u) & ~15u) + ch;
  const uint4 r = philox_block_mask(seed, kind, site, sample, window, e >> 4);
  return keep_at(keep_bits(r, keep_threshold(keep)), e & 15u);
}

__device__ __forceinline__ float philox_sign(uint64_t seed, uint32_t kind, uint32_t site, uint32_t sample,
                                             uint32_t window, uint32_t elem) {
  const uint4 r = philox_block(seed, kind, site, sample, window, elem >> 2);
  return (lane_of(r, elem & 3u) >> 31) ? -1.0f : 1.0f;
}

}  // namespace brl
