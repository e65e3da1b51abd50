// fp32 SIMT engine: implicit-GEMM conv / linear layers with the variational rules fused in the
// gather prologue and the epilogue, plus the small reductions of the ELBO / predictive path.
// This is the parity engine (rtol 1e-3 vs the oracle); the tcgen05 engine lives in brl_tc.cu.
#include <atomic>

#include "brl_kernels.cuh"
#include "brl_philox.cuh"

namespace brl {

std::atomic<long long> g_launch_count{0};
long long launch_count() { return g_launch_count.load(); }
void count_launch(int n) { g_launch_count.fetch_add(n); }

// ------------------------------------------------------------------------------------------------
// helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float noise_normal(const NoiseRef& nz, int s, int b, int B, int per_window, int e) {
  if (nz.ptr) return nz.ptr[((long long)s * B + b) * per_window + e];
  const NoiseKey k = noise_key(nz);
  return philox_normal(k.seed, nz.kind, nz.site, k.sample0 + s, k.window0 + b, e);
}
__device__ __forceinline__ float softplusf(float x) { return x > 20.0f ? x : log1pf(expf(x)); }

__device__ __forceinline__ double block_sum(double v) {
  __shared__ double red[32];
  __syncthreads();
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    v = threadIdx.x < (blockDim.x + 31) / 32 ? red[threadIdx.x] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  }
  return v;  // valid in thread 0
}

// ------------------------------------------------------------------------------------------------
// pooling
// ------------------------------------------------------------------------------------------------
__global__ void maxpool3_kernel(const PoolParams p) {
  const long long total = p.n_img * p.C * p.Hin * p.Win;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int w = i % p.Win;
    const int h = (i / p.Win) % p.Hin;
    const int c = (i / ((long long)p.Win * p.Hin)) % p.C;
    const long long img = i / ((long long)p.Win * p.Hin * p.C);
    const float* src = p.in + img * p.in_img_stride + (long long)c * p.sC + (long long)w * p.sW;
    float v = src[(long long)h * p.sH];
    if (h > 0) v = fmaxf(v, src[(long long)(h - 1) * p.sH]);
    if (h + 1 < p.Hin) v = fmaxf(v, src[(long long)(h + 1) * p.sH]);
    p.out[img * p.out_img_stride + ((long long)c * p.Hin + h) * p.Win + w] = v;
  }
}
void launch_maxpool3(const PoolParams& p, cudaStream_t st) {
  const long long total = p.n_img * p.C * p.Hin * p.Win;
  ++g_launch_count;
  maxpool3_kernel<<<(unsigned)min((total + 255) / 256, (long long)148 * 16), 256, 0, st>>>(p);
}

__device__ __forceinline__ int argmax3_first(const float* src, int o, int H, long long sH) {
  int best = o > 0 ? o - 1 : o;
  float bv = src[(long long)best * sH];
  for (int t = best + 1; t <= min(o + 1, H - 1); ++t) {
    const float v = src[(long long)t * sH];
    if (v > bv) { bv = v; best = t; }
  }
  return best;
}
__global__ void maxpool3_bwd_kernel(const PoolParams p, const float* gout, float* gin) {
  const long long total = p.n_img * p.C * p.Hin * p.Win;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int w = i % p.Win;
    const int h = (i / p.Win) % p.Hin;
    const int c = (i / ((long long)p.Win * p.Hin)) % p.C;
    const long long img = i / ((long long)p.Win * p.Hin * p.C);
    const float* src = p.in + img * p.in_img_stride + (long long)c * p.sC + (long long)w * p.sW;
    const float* go = gout + img * p.out_img_stride + (long long)c * p.Hin * p.Win + w;
    float g = 0.f;
    for (int o = max(h - 1, 0); o <= min(h + 1, p.Hin - 1); ++o)
      if (argmax3_first(src, o, p.Hin, p.sH) == h) g += go[(long long)o * p.Win];
    atomicAdd(gin + img * p.in_img_stride + (long long)c * p.sC + (long long)h * p.sH + (long long)w * p.sW, g);
  }
}
void launch_maxpool3_bwd(const PoolParams& p, const float* gout, float* gin, cudaStream_t st) {
  const long long total = p.n_img * p.C * p.Hin * p.Win;
  ++g_launch_count;
  maxpool3_bwd_kernel<<<(unsigned)min((total + 255) / 256, (long long)148 * 16), 256, 0, st>>>(p, gout, gin);
}

__global__ void avgpool2_kernel(const PoolParams p) {
  const long long total = p.n_img * p.C * p.Hout * p.Win;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int w = i % p.Win;
    const int h = (i / p.Win) % p.Hout;
    const int c = (i / ((long long)p.Win * p.Hout)) % p.C;
    const long long img = i / ((long long)p.Win * p.Hout * p.C);
    const float* src = p.in + img * p.in_img_stride + (long long)c * p.sC + (long long)w * p.sW;
    p.out[img * p.out_img_stride + ((long long)c * p.Hout + h) * p.Win + w] =
        0.5f * (src[(long long)(2 * h) * p.sH] + src[(long long)(2 * h + 1) * p.sH]);
  }
}
void launch_avgpool2(const PoolParams& p, cudaStream_t st) {
  const long long total = p.n_img * p.C * p.Hout * p.Win;
  ++g_launch_count;
  avgpool2_kernel<<<(unsigned)min((total + 255) / 256, (long long)148 * 16), 256, 0, st>>>(p);
}
__global__ void avgpool2_bwd_kernel(const PoolParams p, const float* gout, float* gin) {
  const long long total = p.n_img * p.C * p.Hin * p.Win;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int w = i % p.Win;
    const int h = (i / p.Win) % p.Hin;
    const int c = (i / ((long long)p.Win * p.Hin)) % p.C;
    const long long img = i / ((long long)p.Win * p.Hin * p.C);
    if (h < 2 * p.Hout)
      gin[img * p.in_img_stride + (long long)c * p.sC + (long long)h * p.sH + (long long)w * p.sW] +=
          0.5f * gout[img * p.out_img_stride + ((long long)c * p.Hout + (h >> 1)) * p.Win + w];
  }
}
void launch_avgpool2_bwd(const PoolParams& p, const float* gout, float* gin, cudaStream_t st) {
  const long long total = p.n_img * p.C * p.Hin * p.Win;
  ++g_launch_count;
  avgpool2_bwd_kernel<<<(unsigned)min((total + 255) / 256, (long long)148 * 16), 256, 0, st>>>(p, gout, gin);
}

// ------------------------------------------------------------------------------------------------
// backward through activation / dropout / head
// ------------------------------------------------------------------------------------------------
__global__ void bwd_act_kernel(const BwdAct p) {
  const long long per = (long long)p.N * p.P, total = p.n_img * per;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long img = i / per;
    const int e = (int)(i - img * per);
    const int n = e / p.P, pp = e - n * p.P;
    const long long oidx = img * p.img_stride + (long long)(p.co_off + n) * p.out_P + pp;
    const float g = p.gout[oidx], o = p.outv[oidx];
    float d;
    if (p.head) d = o > 1e-9f ? g * (1.0f - expf(-o)) : 0.f;
    else if (p.relu) d = o > 0.f ? g * p.inv_keep : 0.f;
    else d = g * p.inv_keep;
    p.dpre[i] = d;
    const int s = (int)(img / p.B), b = (int)(img - (long long)s * p.B);
    if (p.dvar) {
      const float sd = p.sd[i];
      const float eps = noise_normal(p.eps, s, b, p.B, (int)per, e);
      p.dvar[i] = sd > 0.f ? d * eps / (2.0f * sd) : 0.f;
    }
    if (p.dpert) p.dpert[i] = d * p.sign_out[img * p.N + n];
  }
}
void launch_bwd_act(const BwdAct& p, cudaStream_t st) {
  const long long total = p.n_img * p.N * p.P;
  ++g_launch_count;
  bwd_act_kernel<<<(unsigned)min((total + 255) / 256, (long long)148 * 16), 256, 0, st>>>(p);
}

// ------------------------------------------------------------------------------------------------
// guide samplers
// ------------------------------------------------------------------------------------------------
__global__ void sample_normal_kernel(const float* __restrict__ mu, const float* __restrict__ sigma, long long P,
                                     NoiseRef eps, float* __restrict__ w, float* __restrict__ delta) {
  const long long nblk = (P + 3) >> 2;
  const int s = blockIdx.y;
  for (long long blk = blockIdx.x * (long long)blockDim.x + threadIdx.x; blk < nblk; blk += (long long)gridDim.x * blockDim.x) {
    float z[4];
    const long long e0 = blk << 2;
    if (eps.ptr) {
#pragma unroll
      for (int l = 0; l < 4; ++l) z[l] = e0 + l < P ? eps.ptr[(long long)s * P + e0 + l] : 0.f;
    } else {
      const NoiseKey k = noise_key(eps);
      const float4 n4 = normal4(philox_block(k.seed, KIND_WEIGHT_EPS, 0, k.sample0 + s, 0, (uint32_t)blk));
      z[0] = n4.x; z[1] = n4.y; z[2] = n4.z; z[3] = n4.w;
    }
#pragma unroll
    for (int l = 0; l < 4; ++l) {
      const long long e = e0 + l;
      if (e < P) {
        w[(long long)s * P + e] = fmaf(sigma[e], z[l], mu[e]);
        if (delta) delta[(long long)s * P + e] = z[l];
      }
    }
  }
}
void launch_sample_normal(const float* mu, const float* sigma, long long P, long long S, NoiseRef eps, float* w,
                          float* delta, cudaStream_t st) {
  const long long nblk = (P + 3) >> 2;
  dim3 grid((unsigned)min((nblk + 255) / 256, (long long)148 * 8), (unsigned)S);
  ++g_launch_count;
  sample_normal_kernel<<<grid, 256, 0, st>>>(mu, sigma, P, eps, w, delta);
}

// The radial guide needs ||eps|| over each whole site before any weight of the site exists, so the draw is two passes over
// the same Philox stream (the stream of sample_normal_kernel: block = element >> 2, four normals per block).
__device__ __forceinline__ void weight_eps4(const NoiseRef& eps, const NoiseKey& k, int s, long long P, long long blk, float (&z)[4]) {
  const long long e0 = blk << 2;
  if (eps.ptr) {
#pragma unroll
    for (int l = 0; l < 4; ++l) z[l] = e0 + l < P ? eps.ptr[(long long)s * P + e0 + l] : 0.f;
  } else {
    const float4 n4 = normal4(philox_block(k.seed, KIND_WEIGHT_EPS, 0, k.sample0 + s, 0, (uint32_t)blk));
    z[0] = n4.x; z[1] = n4.y; z[2] = n4.z; z[3] = n4.w;
  }
}
__global__ void radial_norm_kernel(long long P, const long long* __restrict__ site_off, int n_sites, NoiseRef eps,
                                   float* __restrict__ norms) {
  const int j = blockIdx.x, s = blockIdx.y;
  const long long beg = site_off[j], end = site_off[j + 1];
  const NoiseKey k = noise_key(eps);
  double acc = 0.0;
  for (long long blk = (beg >> 2) + threadIdx.x; blk <= ((end - 1) >> 2); blk += blockDim.x) {
    float z[4];
    weight_eps4(eps, k, s, P, blk, z);
#pragma unroll
    for (int l = 0; l < 4; ++l) {
      const long long e = (blk << 2) + l;
      if (e >= beg && e < end) acc += (double)z[l] * z[l];
    }
  }
  acc = block_sum(acc);
  if (threadIdx.x == 0) norms[(long long)s * n_sites + j] = (float)sqrt(acc);
}
__global__ void radial_apply_kernel(const float* __restrict__ mu, const float* __restrict__ sigma, long long P,
                                    const long long* __restrict__ site_off, int n_sites, NoiseRef eps, NoiseRef r,
                                    const float* __restrict__ norms, float* __restrict__ w, float* __restrict__ delta) {
  const int s = blockIdx.y;
  const long long nblk = (P + 3) >> 2;
  const NoiseKey k = noise_key(eps), rk = noise_key(r);
  for (long long blk = blockIdx.x * (long long)blockDim.x + threadIdx.x; blk < nblk; blk += (long long)gridDim.x * blockDim.x) {
    float z[4];
    weight_eps4(eps, k, s, P, blk, z);
    const long long e0 = blk << 2;
    int lo = 0, hi = n_sites - 1;  // site of e0: last j with site_off[j] <= e0
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (site_off[mid] <= e0) lo = mid; else hi = mid - 1;
    }
    int j = lo;
    float scale = 0.f;
    bool have = false;
#pragma unroll
    for (int l = 0; l < 4; ++l) {
      const long long e = e0 + l;
      if (e >= P) break;
      while (e >= site_off[j + 1]) { ++j; have = false; }
      if (!have) {
        const float rr = r.ptr ? r.ptr[(long long)s * n_sites + j]
                               : philox_normal(rk.seed, KIND_RADIAL_R, 0, rk.sample0 + s, 0, (uint32_t)j);
        scale = rr / norms[(long long)s * n_sites + j];
        have = true;
      }
      const float d = z[l] * scale;
      w[(long long)s * P + e] = fmaf(d, sigma[e], mu[e]);
      if (delta) delta[(long long)s * P + e] = d;
    }
  }
}
void launch_sample_radial(const float* mu, const float* sigma, long long P, long long S, const long long* site_off,
                          int n_sites, int max_site, NoiseRef eps, NoiseRef r, float* norms, float* w, float* delta,
                          cudaStream_t st) {
  ++g_launch_count;
  radial_norm_kernel<<<dim3(n_sites, (unsigned)S), 256, 0, st>>>(P, site_off, n_sites, eps, norms);
  ++g_launch_count;
  const long long nblk = (P + 3) >> 2;
  radial_apply_kernel<<<dim3((unsigned)min((nblk + 255) / 256, (long long)148 * 8), (unsigned)S), 256, 0, st>>>(
      mu, sigma, P, site_off, n_sites, eps, r, norms, w, delta);
}

__global__ void gen_signs_kernel(float* dst, long long S, long long B, int C, NoiseRef nz) {
  const long long total = S * B * C;
  const NoiseKey k = noise_key(nz);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = i % C;
    const long long sb = i / C;
    const int b = sb % B, s = sb / B;
    dst[i] = philox_sign(k.seed, nz.kind, nz.site, k.sample0 + s, k.window0 + b, c);
  }
}
__global__ void gen_signs_multi_kernel(const SignJobs jobs, long long B, NoiseRef base) {
  const long long total = jobs.start[jobs.n];
  const NoiseKey k = noise_key(base);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int j = 0;
    while (j + 1 < jobs.n && i >= jobs.start[j + 1]) ++j;
    // one work item = one Philox block = the signs of four consecutive channels of a window (element c: block c >> 2, lane c & 3,
    // the keying of philox_sign)
    const long long e = i - jobs.start[j];
    const int C = jobs.C[j], Q = (C + 3) >> 2, q = (int)(e % Q), b = (int)(e / Q);
    const uint4 r = philox_block(k.seed, jobs.kind[j], jobs.site[j], k.sample0, k.window0 + b, (uint32_t)q);
    float* d = jobs.dst[j] + (long long)b * C + 4 * q;
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int l = 0; l < 4; ++l)
      if (4 * q + l < C) d[l] = (w[l] >> 31) ? -1.0f : 1.0f;
  }
}
void launch_gen_signs_multi(const SignJobs& jobs, long long B, NoiseRef base, cudaStream_t st) {
  const long long total = jobs.start[jobs.n];
  if (total <= 0) return;
  ++g_launch_count;
  gen_signs_multi_kernel<<<(unsigned)min((total + 255) / 256, (long long)148 * 8), 256, 0, st>>>(jobs, B, base);
}
void launch_gen_signs(float* dst, long long S, long long B, int C, NoiseRef nz, cudaStream_t st) {
  const long long total = S * B * C;
  ++g_launch_count;
  gen_signs_kernel<<<(unsigned)min((total + 255) / 256, (long long)148 * 8), 256, 0, st>>>(dst, S, B, C, nz);
}

// ------------------------------------------------------------------------------------------------
// likelihoods
// ------------------------------------------------------------------------------------------------
__global__ void nll_elbo_kernel(const float* __restrict__ out, const float* __restrict__ y, long long B, float gscale,
                                double* acc, float* __restrict__ gout) {
  double nll = 0.0, se = 0.0;
  for (long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x; b < B; b += (long long)gridDim.x * blockDim.x) {
    const float loc = out[2 * b], sc = out[2 * b + 1];
    const float s = softplusf(sc);  // second softplus: positive_scale=False (bayesian.py:73-76)
    const float d = y[b] - loc, r = d / s;
    nll += 0.5 * (double)r * r + (double)logf(s) + 0.9189385332046727;
    se += (double)d * d;
    if (gout) {
      const float sig = sc > 20.0f ? 1.0f : 1.0f / (1.0f + expf(-sc));
      gout[2 * b] = -r / s * gscale;
      gout[2 * b + 1] = (1.0f - r * r) / s * sig * gscale;
    }
  }
  nll = block_sum(nll);
  se = block_sum(se);
  if (threadIdx.x == 0) {
    atomicAdd(acc + 0, nll);
    atomicAdd(acc + 1, se);
  }
}
void launch_nll_elbo(const float* out, const float* y, long long B, float gscale, double* acc, float* gout,
                     cudaStream_t st) {
  ++g_launch_count;
  nll_elbo_kernel<<<(unsigned)min((B + 255) / 256, (long long)148), 256, 0, st>>>(out, y, B, gscale, acc, gout);
}

__global__ void nll_hnn_kernel(const float* __restrict__ out, const float* __restrict__ y, long long B, double* acc,
                               float* __restrict__ gout) {
  double loss = 0.0, se = 0.0;
  const float invB = 1.0f / (float)B;
  for (long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x; b < B; b += (long long)gridDim.x * blockDim.x) {
    const float loc = out[2 * b], sc = out[2 * b + 1];
    const float var0 = sc * sc, var = fmaxf(var0, 1e-6f);  // F.gaussian_nll_loss eps clamp
    const float d = loc - y[b];
    loss += 0.5 * ((double)logf(var) + (double)d * d / var);
    se += (double)d * d;
    if (gout) {
      gout[2 * b] = d / var * invB;
      gout[2 * b + 1] = var0 > 1e-6f ? 0.5f * (1.0f / var - d * d / (var * var)) * 2.0f * sc * invB : 0.f;
    }
  }
  loss = block_sum(loss);
  se = block_sum(se);
  if (threadIdx.x == 0) {
    atomicAdd(acc + 0, loss / (double)B);
    atomicAdd(acc + 1, se / (double)B);
  }
}
void launch_nll_hnn(const float* out, const float* y, long long B, double* acc, float* gout, cudaStream_t st) {
  ++g_launch_count;
  nll_hnn_kernel<<<(unsigned)min((B + 255) / 256, (long long)148), 256, 0, st>>>(out, y, B, acc, gout);
}

// ------------------------------------------------------------------------------------------------
// KL + gradient finalisation
// ------------------------------------------------------------------------------------------------
__global__ void finalize_kernel(const Finalize p) {
  double kl = 0.0;
  const float isp2 = 1.0f / (p.prior_scale * p.prior_scale);
  const float lsp = logf(p.prior_scale);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < p.P; i += (long long)gridDim.x * blockDim.x) {
    const float mu = p.mu[i], sg = p.sigma[i];
    float gmu = 0.f, gsg = 0.f, dmu_kl, dsg_kl;
    if (p.guide == 1) {  // radial: sampled log q - log p (Trace_ELBO)
      const float w = p.w[i], d = p.delta[i];
      const float zp = (w - p.prior_loc) / p.prior_scale;
      kl += (double)(-logf(sg) - 0.5f * d * d) - (double)(-lsp - 0.5f * zp * zp);
      dmu_kl = (w - p.prior_loc) * isp2;
      dsg_kl = -1.0f / sg + (w - p.prior_loc) * isp2 * d;
      if (p.g0) { gmu = p.g0[i]; gsg = p.g0[i] * d; }
    } else {
      const float dm = mu - p.prior_loc;
      kl += (double)(lsp - logf(sg)) + (double)((sg * sg + dm * dm) * 0.5f * isp2) - 0.5;
      dmu_kl = dm * isp2;
      dsg_kl = -1.0f / sg + sg * isp2;
      if (p.g0) {
        gmu = p.g0[i];
        if (p.mode == 2) gsg = 2.0f * sg * p.g1[i];       // LRT: d/dsigma through sigma^2
        else if (p.mode == 3) gsg = p.g1[i] * p.delta[i]; // flipout: dDeltaW * eps
        else gsg = p.g0[i] * p.delta[i];                   // weight sampling
      }
    }
    if (p.grad_mu) {
      gmu = fmaf(p.c_kl, dmu_kl, gmu);
      gsg = fmaf(p.c_kl, dsg_kl, gsg);
      if (!p.first) { gmu += p.grad_mu[i]; gsg += p.grad_sigma[i]; }
      p.grad_mu[i] = gmu;
      p.grad_sigma[i] = gsg;
      if (p.grad_log_sigma) p.grad_log_sigma[i] = gsg * sg;
    }
  }
  kl = block_sum(kl);
  if (threadIdx.x == 0) atomicAdd(p.kl_acc, kl);
}
void launch_finalize(const Finalize& p, cudaStream_t st) {
  ++g_launch_count;
  finalize_kernel<<<(unsigned)min((p.P + 255) / 256, (long long)148 * 4), 256, 0, st>>>(p);
}

__global__ void post_scalars_kernel(double* scalars, const double* acc, double c_nll, double c, int particles,
                                    long long B) {
  // acc = {nll_sum over particles, squared error sum over particles, kl sum over particles}
  scalars[0] = (c_nll * acc[0] + c * acc[2]) / particles;
  scalars[1] = acc[0] / particles;
  scalars[2] = acc[2] / particles;
  scalars[3] = acc[1] / ((double)particles * (double)B);
}
void launch_post_scalars(double* scalars, const double* acc, double c_nll, double c, int particles, long long B,
                         cudaStream_t st) {
  ++g_launch_count;
  post_scalars_kernel<<<1, 1, 0, st>>>(scalars, acc, c_nll, c, particles, B);
}

__global__ void log_sigma_grad_kernel(const float* gs, const float* sigma, float* gls, long long P) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < P; i += (long long)gridDim.x * blockDim.x)
    gls[i] = gs[i] * sigma[i];
}
void launch_log_sigma_grad(const float* gs, const float* sigma, float* gls, long long P, cudaStream_t st) {
  ++g_launch_count;
  log_sigma_grad_kernel<<<(unsigned)min((P + 255) / 256, (long long)148 * 4), 256, 0, st>>>(gs, sigma, gls, P);
}

// ------------------------------------------------------------------------------------------------
// predictive moments (Welford over MC samples; chunks merge through `state`)
// ------------------------------------------------------------------------------------------------
__global__ void moments_update_kernel(const float* __restrict__ out, long long S, long long B, float* state, int first) {
  for (long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x; b < B; b += (long long)gridDim.x * blockDim.x) {
    float n = 0.f, mean = 0.f, m2 = 0.f, s2 = 0.f;
    if (!first) { n = state[b]; mean = state[B + b]; m2 = state[2 * B + b]; s2 = state[3 * B + b]; }
    // the loads do not depend on the Welford chain: batches of 16 are issued ahead of it (the loop was one DRAM round trip per MC
    // sample: 52 us for S = 100); the arithmetic and its order are unchanged
    long long s = 0;
    for (; s + 16 <= S; s += 16) {
      float2 o[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) o[u] = __ldcs(reinterpret_cast<const float2*>(out + ((s + u) * B + b) * 2));
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        n += 1.f;
        const float d = o[u].x - mean;
        mean += d / n;
        m2 = fmaf(d, o[u].x - mean, m2);
        s2 = fmaf(o[u].y, o[u].y, s2);
      }
    }
    for (; s < S; ++s) {
      const float2 o = *reinterpret_cast<const float2*>(out + (s * B + b) * 2);
      n += 1.f;
      const float d = o.x - mean;
      mean += d / n;
      m2 = fmaf(d, o.x - mean, m2);
      s2 = fmaf(o.y, o.y, s2);
    }
    state[b] = n; state[B + b] = mean; state[2 * B + b] = m2; state[3 * B + b] = s2;
  }
}
void launch_moments_update(const float* out, long long S, long long B, float* state, int first, cudaStream_t st) {
  ++g_launch_count;
  moments_update_kernel<<<(unsigned)min((B + 127) / 128, (long long)148 * 8), 128, 0, st>>>(out, S, B, state, first);
}
__global__ void moments_final_kernel(const float* __restrict__ state, long long B, float* pred, float* std, float* ep,
                                     float* al) {
  for (long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x; b < B; b += (long long)gridDim.x * blockDim.x) {
    const float n = state[b];
    const float e = state[2 * B + b] / (n - 1.f);  // unbiased: loc.var(0) (bayesian.py:212)
    const float a = state[3 * B + b] / n;
    pred[b] = state[B + b];
    if (ep) ep[b] = e;
    if (al) al[b] = a;
    std[b] = sqrtf(a + e);
  }
}
void launch_moments_final(const float* state, long long B, float* pred, float* std, float* ep, float* al,
                          cudaStream_t st) {
  ++g_launch_count;
  moments_final_kernel<<<(unsigned)min((B + 255) / 256, (long long)148 * 8), 256, 0, st>>>(state, B, pred, std, ep, al);
}

__global__ void moments_direct_kernel(const float* __restrict__ out, long long S, long long B, float* pred, float* std,
                                      float* ep, float* al) {
  for (long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x; b < B; b += (long long)gridDim.x * blockDim.x) {
    float n = 0.f, mean = 0.f, m2 = 0.f, s2 = 0.f;
    for (long long s = 0; s < S; ++s) {
      const float2 o = *reinterpret_cast<const float2*>(out + (s * B + b) * 2);
      n += 1.f;
      const float d = o.x - mean;
      mean += d / n;
      m2 = fmaf(d, o.x - mean, m2);
      s2 = fmaf(o.y, o.y, s2);
    }
    const float e = m2 / (n - 1.f), a = s2 / n;
    pred[b] = mean; ep[b] = e; al[b] = a; std[b] = sqrtf(a + e);
  }
}
void launch_moments_direct(const float* out, long long S, long long B, float* pred, float* std, float* ep, float* al,
                           cudaStream_t st) {
  ++g_launch_count;
  moments_direct_kernel<<<(unsigned)min((B + 127) / 128, (long long)148 * 8), 128, 0, st>>>(out, S, B, pred, std, ep, al);
}

__global__ void aggregate_kernel(const float* __restrict__ out, long long S, long long B, float* agg) {
  for (long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x; b < B; b += (long long)gridDim.x * blockDim.x) {
    float n = 0.f, mean = 0.f, m2 = 0.f, s2 = 0.f, wl = 0.f, wp = 0.f;
    for (long long s = 0; s < S; ++s) {
      const float2 o = *reinterpret_cast<const float2*>(out + (s * B + b) * 2);
      n += 1.f;
      const float d = o.x - mean;
      mean += d / n;
      m2 = fmaf(d, o.x - mean, m2);
      s2 = fmaf(o.y, o.y, s2);
      const float prec = 1.0f / (o.y * o.y);
      wl = fmaf(o.x, prec, wl);
      wp += prec;
    }
    const float sc = sqrtf(s2 / n + m2 / (n - 1.f));
    agg[2 * b] = wl / wp;
    agg[2 * b + 1] = sc + logf(-expm1f(-sc));  // inverse softplus (positive_scale=False)
  }
}
void launch_aggregate(const float* out, long long S, long long B, float* agg, cudaStream_t st) {
  ++g_launch_count;
  aggregate_kernel<<<(unsigned)min((B + 127) / 128, (long long)148 * 8), 128, 0, st>>>(out, S, B, agg);
}

__global__ void mixture_kernel(const float* __restrict__ mu_m, const float* __restrict__ sd_m, long long M, long long n,
                               float* mu, float* sd) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    double a = 0.0, q = 0.0;
    for (long long m = 0; m < M; ++m) {
      const double u = mu_m[m * n + i], s = sd_m[m * n + i];
      a += u;
      q += u * u + s * s;
    }
    a /= (double)M;
    mu[i] = (float)a;
    sd[i] = (float)sqrt(q / (double)M - a * a);  // biased mixture variance (deepens.py:24)
  }
}
void launch_mixture(const float* mu_m, const float* sd_m, long long M, long long n, float* mu, float* sd,
                    cudaStream_t st) {
  ++g_launch_count;
  mixture_kernel<<<(unsigned)min((n + 255) / 256, (long long)148 * 8), 256, 0, st>>>(mu_m, sd_m, M, n, mu, sd);
}

// ------------------------------------------------------------------------------------------------
// test-time scalar metrics (bayesian.py:217-224; results/metrics.py:210-274)
// ------------------------------------------------------------------------------------------------
__global__ void test_metrics_acc_kernel(const float* __restrict__ pred, const float* __restrict__ std,
                                        const float* __restrict__ y, long long n, double* acc, unsigned int* hist) {
  __shared__ unsigned int sh[100];
  for (int i = threadIdx.x; i < 100; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  double nll = 0.0, se = 0.0, s2 = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float sd = std[i], d = pred[i] - y[i];
    const float var = fmaxf(sd * sd, 1e-6f);
    nll += 0.5 * ((double)logf(var) + (double)d * d / var);
    se += (double)d * d;
    s2 += (double)sd * sd;
    // interval coverage: |z| <= icdf(0.5 + p/2)  <=>  erf(|z|/sqrt2) <= p ; first bin j with p_j >= q
    const float q = erff(fabsf(d / sd) * 0.70710678f);
    int j = (int)ceilf(q * 99.0f - 1e-6f);
    j = max(0, min(99, j));
    atomicAdd(&sh[j], 1u);
  }
  nll = block_sum(nll);
  se = block_sum(se);
  s2 = block_sum(s2);
  if (threadIdx.x == 0) { atomicAdd(acc + 0, nll); atomicAdd(acc + 1, se); atomicAdd(acc + 2, s2); }
  __syncthreads();
  for (int i = threadIdx.x; i < 100; i += blockDim.x)
    if (sh[i]) atomicAdd(hist + i, sh[i]);
}
__global__ void test_metrics_final_kernel(const double* acc, const unsigned int* hist, long long n, double* scalars, int nscal) {
  if (threadIdx.x != 0) return;
  scalars[0] = acc[0] / (double)n;
  scalars[1] = acc[1] / (double)n;
  scalars[2] = sqrt(acc[2] / (double)n);
  double cum = 0.0, sq = 0.0, ab = 0.0;
  for (int j = 0; j < 100; ++j) {
    cum += hist[j];
    const double e = (double)j / 99.0 - cum / (double)n;
    sq += e * e;
    ab += fabs(e);
  }
  scalars[3] = sqrt(sq / 100.0);
  if (nscal > 4) scalars[4] = ab / 100.0;  // mean absolute calibration error (results/metrics.py:277-297)
}
void launch_test_metrics(const float* pred, const float* std, const float* y, long long n, double* scalars,
                         unsigned int* hist, cudaStream_t st, int nscal) {
  // workspace layout: hist[100] u32 followed (at +512 B) by acc[3] doubles
  double* acc = reinterpret_cast<double*>(reinterpret_cast<char*>(hist) + 512);
  cudaMemsetAsync(hist, 0, 512 + 3 * sizeof(double), st);
  ++g_launch_count;
  test_metrics_acc_kernel<<<(unsigned)min((n + 255) / 256, (long long)148 * 4), 256, 0, st>>>(pred, std, y, n, acc, hist);
  ++g_launch_count;
  test_metrics_final_kernel<<<1, 32, 0, st>>>(acc, hist, n, scalars, nscal);
}

// ------------------------------------------------------------------------------------------------
// ClippedAdam
// ------------------------------------------------------------------------------------------------
__global__ void clipped_adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                    float* __restrict__ v, long long n, float step_size, float b1, float b2, float eps,
                                    float clip, float wd) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float gr = fminf(fmaxf(g[i], -clip), clip);
    const float pi = p[i];
    if (wd != 0.f) gr = fmaf(wd, pi, gr);
    const float mi = b1 * m[i] + (1.f - b1) * gr;
    const float vi = b2 * v[i] + (1.f - b2) * gr * gr;
    m[i] = mi;
    v[i] = vi;
    p[i] = pi - step_size * mi / (sqrtf(vi) + eps);
  }
}
// the optimiser step of a mean-field guide in one launch: ClippedAdam on loc and on log scale (the unconstrained parameter of
// Pyro's positive constraint), then scale = exp(log scale) for the next forward pass
__global__ void clipped_adam_vi_kernel(float* __restrict__ loc, float* __restrict__ ls, float* __restrict__ scale,
                                       const float* __restrict__ g_loc, const float* __restrict__ g_ls, float* __restrict__ m_loc,
                                       float* __restrict__ v_loc, float* __restrict__ m_ls, float* __restrict__ v_ls, long long n,
                                       float step_size, float b1, float b2, float eps, float clip, float wd, float gscale) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    {
      float gr = fminf(fmaxf(g_loc[i] * gscale, -clip), clip);
      const float pi = loc[i];
      if (wd != 0.f) gr = fmaf(wd, pi, gr);
      const float mi = b1 * m_loc[i] + (1.f - b1) * gr, vi = b2 * v_loc[i] + (1.f - b2) * gr * gr;
      m_loc[i] = mi;
      v_loc[i] = vi;
      loc[i] = pi - step_size * mi / (sqrtf(vi) + eps);
    }
    {
      float gr = fminf(fmaxf(g_ls[i] * gscale, -clip), clip);
      const float pi = ls[i];
      if (wd != 0.f) gr = fmaf(wd, pi, gr);
      const float mi = b1 * m_ls[i] + (1.f - b1) * gr, vi = b2 * v_ls[i] + (1.f - b2) * gr * gr;
      m_ls[i] = mi;
      v_ls[i] = vi;
      const float pn = pi - step_size * mi / (sqrtf(vi) + eps);
      ls[i] = pn;
      scale[i] = expf(pn);
    }
  }
}
void launch_clipped_adam_vi(float* loc, float* ls, float* scale, const float* g_loc, const float* g_ls, float* m_loc, float* v_loc,
                            float* m_ls, float* v_ls, long long n, float step_size, float b1, float b2, float eps, float clip, float wd,
                            cudaStream_t st, float gscale) {
  ++g_launch_count;
  clipped_adam_vi_kernel<<<(unsigned)min((n + 255) / 256, (long long)148 * 8), 256, 0, st>>>(loc, ls, scale, g_loc, g_ls, m_loc, v_loc,
                                                                                            m_ls, v_ls, n, step_size, b1, b2, eps, clip, wd,
                                                                                            gscale);
}
void launch_clipped_adam(float* p, const float* g, float* m, float* v, long long n, float step_size, float b1,
                         float b2, float eps, float clip, float wd, cudaStream_t st) {
  ++g_launch_count;
  clipped_adam_kernel<<<(unsigned)min((n + 255) / 256, (long long)148 * 8), 256, 0, st>>>(p, g, m, v, n, step_size, b1, b2,
                                                                                         eps, clip, wd);
}

}  // namespace brl
