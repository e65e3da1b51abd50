// Fused epilogues of the forward / input-gradient implicit GEMMs, shared by the fp32 SIMT kernels (brl_gemm.cu) and the
// tcgen05 TF32 kernels (brl_tc_gemm.cu): bias, LRT eps * sqrt(var), Flipout sign flips, ReLU, dropout, softplus head.
#pragma once
#include "brl_kernels.cuh"
#include "brl_philox.cuh"

namespace brl {

__device__ __forceinline__ float gnoise_normal(const NoiseRef& nz, int s, int b, int B, int per_window, int e) {
  if (nz.ptr) return nz.ptr[((long long)s * B + b) * per_window + e];
  const NoiseKey k = noise_key(nz);
  return philox_normal(k.seed, nz.kind, nz.site, k.sample0 + s, k.window0 + b, e);
}
// dropout site of an op with C channels x P positions: injected masks are [S,B,C,P] (reference layout)
__device__ __forceinline__ bool gnoise_keep(const NoiseRef& nz, int s, int b, int B, int C, int P, int ch, int pos, float keep) {
  if (nz.ptr) return nz.ptr[(((long long)s * B + b) * C + ch) * P + pos] != 0.0f;
  const NoiseKey k = noise_key(nz);
  return philox_keep(k.seed, nz.kind, nz.site, k.sample0 + s, k.window0 + b, C, pos, ch, keep);
}

// one output element of a forward / input-gradient GEMM
template <int EPI>
__device__ __forceinline__ void gemm_epilogue(const ConvGemm& p, int s, int m, int n, float a0, float a1) {
  const int b = m / p.P, pp = m - b * p.P;
  const long long img = (long long)s * p.B + b;
  const long long oidx = img * p.out_img_stride + (long long)(p.co_off + n) * p.out_P + pp;
  if (EPI >= EPI_DX_PLAIN) {
    float g = a0;
    if (EPI == EPI_DX_LRT) g = fmaf(2.0f * p.xin[oidx], a1, g);
    if (EPI == EPI_DX_FLIPOUT) g = fmaf(p.sign_in[img * p.sign_C + n], a1, g);
    atomicAdd(p.out + oidx, g);  // ops of one backward level accumulate into a shared input-gradient buffer concurrently
    return;
  }
  float v = a0;
  if (EPI == EPI_FWD_PLAIN) {
    v += p.bias0[(long long)s * p.bs0 + n];
  } else if (EPI == EPI_FWD_LRT) {
    const float mean = v + p.bias0[n];
    const float sb = p.bias1[n];
    float var = fmaf(sb, sb, a1);
    if (var < 0.f) var += fabsf(var) + 1e-6f;
    const float sd = sqrtf(var);
    const float e = gnoise_normal(p.eps, s, b, p.B, p.N * p.P, n * p.P + pp);
    v = fmaf(sd, e, mean);
    if (p.sd_out) p.sd_out[(img * p.N + n) * p.P + pp] = sd;
  } else {  // flipout
    v = v + a1 * p.sign_out[img * p.N + n] + p.bias1[(long long)s * p.bs1 + n];
  }
  if (p.relu) v = fmaxf(v, 0.f);
  if (p.keep < 1.0f) v = gnoise_keep(p.drop, s, b, p.B, p.N, p.P, n, pp, p.keep) ? v / p.keep : 0.f;
  if (p.head) {
    v = v > 20.0f ? v : log1pf(expf(v));
    v = v > 1e-9f ? v : 1e-9f;
  }
  p.out[oidx] = v;
}

}  // namespace brl
