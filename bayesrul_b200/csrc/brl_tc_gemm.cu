// tcgen05 / TMEM back-end of the per-layer implicit GEMMs (TF32 operands, fp32 accumulation in TMEM).
//
// Same operator descriptors (ConvGemm / ConvDw) and the same fused epilogues as the fp32 SIMT kernels of brl_gemm.cu;
// only the contraction moves from FFMA to the tensor pipe.  One layer = one kernel:
//   tc_gemm_kernel<DUAL, EPI>  forward / input-gradient pass.  LRT runs x * mu and x^2 * sigma^2 as a DUAL GEMM into two
//                              TMEM accumulators and the epilogue applies eps * sqrt(var); Flipout gathers (x . s_in) as
//                              the second A operand against (W - mu) and applies s_out in the epilogue.
//   tc_dw_kernel               weight gradient  dW[k][co] = sum_m tr(A[m][k]) * G[m][co]  (rows m are the MMA K dimension)
// A CTA owns a 128-row tile.  All 256 threads gather the implicit-GEMM operands from fp32 global memory (im2col tables,
// squares / sign flips applied on the fly, cvt.rna.tf32), write them in the UMMA canonical K-major no-swizzle layout
// ([4-element k chunk][128 rows][16 B]) into a 2-stage shared-memory ring, and one elected lane of warp 0 issues
// kind::tf32 MMAs (M = 128, N = 16..128, K = 8) whose commits free the stages.  The accumulators are read back with
// tcgen05.ld, parked in shared memory (the ring is dead by then) and walked by the rolled epilogue of the SIMT path.
// Stated accuracy: TF32 (10-bit mantissa) products, fp32 accumulation -- parity tests use rtol 5e-3 vs the fp32 engine.
#include <algorithm>
#include <atomic>

#include "brl_gemm_epi.cuh"
#include "brl_kernels.cuh"
#include "brl_philox.cuh"
#include "brl_tc_ptx.cuh"

namespace brl {

extern std::atomic<long long> g_launch_count;

constexpr int TG_K = 32;                       // k elements per stage
constexpr int TG_CHUNK = 128 * 16;             // bytes of one 4-element k chunk of a 128-row operand tile
constexpr int TG_TILE = (TG_K / 4) * TG_CHUNK; // 16 KB
constexpr int TG_STAGE = 4 * TG_TILE;          // A0 | A1 | B0 | B1
constexpr int TG_SMEM = 2 * TG_STAGE + 64;     // + mbarriers / tmem slot

// kind::tf32 instruction descriptor: TF32 A/B (format 2), fp32 D, both K-major, M = 128
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ float to_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}

// common prologue / MMA plumbing of the two kernels
struct TgCtx {
  uint32_t sbase, bars, tmem;
  int warp, lane;
};
__device__ __forceinline__ TgCtx tg_setup(unsigned char* smem, int ncols) {
  TgCtx c;
  c.sbase = smem_u32(smem);
  c.bars = c.sbase + 2 * TG_STAGE;
  c.lane = threadIdx.x & 31;
  c.warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  uint32_t* slot = reinterpret_cast<uint32_t*>(smem + 2 * TG_STAGE + 32);
  if (threadIdx.x == 0) {
    mbar_init(c.bars, 1);
    mbar_init(c.bars + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (c.warp == 0) tmem_alloc(smem_u32(slot), ncols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  c.tmem = *slot;
  return c;
}
__device__ __forceinline__ void tg_teardown(const TgCtx& c, int ncols) {
  tc_fence_before();
  __syncthreads();
  if (c.warp == 0) tmem_dealloc(c.tmem, ncols);
}

// ------------------------------------------------------------------------------------------------
// forward / input-gradient
// ------------------------------------------------------------------------------------------------
template <bool DUAL, int EPI>
__global__ void __launch_bounds__(256, 1) tc_gemm_kernel(const ConvGemm p, int* status) {
  extern __shared__ __align__(128) unsigned char smem[];
  const TgCtx c = tg_setup(smem, 256);
  const int tid = threadIdx.x;
  const bool split = p.ksplit > 1;
  const int s = split ? 0 : blockIdx.z;
  const int m0 = blockIdx.x * 128, n0 = blockIdx.y * 128;
  const int Mtot = p.B * p.P;
  const int ncol = min(128, p.N - n0), NP = (ncol + 15) & ~15;
  int kbeg = 0, kend = p.K;
  if (split) {
    const int per = ((p.K + p.ksplit - 1) / p.ksplit + TG_K - 1) / TG_K * TG_K;
    kbeg = blockIdx.z * per;
    kend = min(p.K, kbeg + per);
  }
  const int nchunk = kend > kbeg ? (kend - kbeg + TG_K - 1) / TG_K : 0;

  // gather roles: thread = (row / column r of the tile, half kq of the stage's k range)
  const int r = tid & 127, kq = tid >> 7;
  const int am = m0 + r;
  const bool arv = am < Mtot;
  const int ab = arv ? am / p.P : 0;
  const int app = arv ? am - ab * p.P : 0;
  const int aoh = app / p.Wrow, aow = app - aoh * p.Wrow;
  const long long aimg = p.a.per_sample ? (long long)s * p.B + ab : ab;
  const long long rowbase = aimg * p.a.img_stride + (long long)aoh * p.a.sH + (long long)aow * p.a.sW;
  const float* sgn = p.sign_in + ((long long)s * p.B + ab) * p.sign_C;
  const float* W0 = p.W0 + (long long)s * p.ws0;
  const float* W1 = DUAL ? p.W1 + (long long)s * p.ws1 : nullptr;

  for (int ch = 0; ch < nchunk; ++ch) {
    const int st = ch & 1;
    unsigned char* stage = smem + st * TG_STAGE;
    if (ch >= 2) {  // the MMAs that read this stage two chunks ago have completed
      mbar_wait(c.bars + 8 * st, ((ch >> 1) - 1) & 1, status, 20);
      tc_fence_after();
    }
    const int k0 = kbeg + ch * TG_K + kq * 16;
    // three passes so that every load of the stage is in flight at once (the stage is latency-, not bandwidth-bound):
    // 1. gather tables, 2. operand loads, 3. transforms + TF32 rounding + 16-byte stores
    int ko[16], kcl[16];
    bool oka[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int k = k0 + j;
      kcl[j] = min(k, p.K - 1);
      const int dhw = __ldg(p.a.kdhw + kcl[j]);
      ko[j] = __ldg(p.a.koff + kcl[j]);
      const int ih = aoh + (int)(short)(dhw & 0xffff), iw = aow + (dhw >> 16);
      oka[j] = arv && k < kend && (unsigned)ih < (unsigned)p.a.Hin && (unsigned)iw < (unsigned)p.a.Win;
    }
    float va[16], va1[16];
    const bool two_bases = DUAL && p.a.base1 != p.a.base0;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const long long off = oka[j] ? rowbase + ko[j] : 0;
      va[j] = __ldg(p.a.base0 + off);
      if (DUAL) {
        if (two_bases) va1[j] = __ldg(p.a.base1 + off);
        else if (p.trA == TRA_SIGN) va1[j] = __ldg(sgn + (oka[j] ? __ldg(p.a.kci + kcl[j]) : 0));
      }
    }
    // B operand (weights): NP rows x 8 four-element k chunks = NP * 8 sixteen-byte units, spread over the 256 threads,
    // so a layer with 16 output channels gathers 1/8 of what a 128-channel tile does
    const int kb0 = kbeg + ch * TG_K;
    for (int u = tid; u < NP * 8; u += 256) {
      const int n = u >> 3, kc4 = u & 7;
      const bool nv = n < ncol;
      const long long bcol_u = (long long)(nv ? n0 + n : 0) * p.nB;
      float w0[4], w1[4];
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const int k = kb0 + 4 * kc4 + jj;
        const int kc = min(k, p.K - 1);
        const long long boff = (p.kB ? (long long)__ldg(p.kB + kc) : (long long)kc) + bcol_u;
        const bool okb = nv && k < kend;
        const float wv = __ldg(W0 + boff);
        w0[jj] = okb ? to_tf32(wv) : 0.f;
        if (DUAL) {
          float w = __ldg(W1 + boff);
          if (p.trB == TRB_SQUARE) w = w * w;
          else if (p.trB == TRB_MINUS_W0) w -= wv;
          w1[jj] = okb ? to_tf32(w) : 0.f;
        }
      }
      const int off = kc4 * TG_CHUNK + n * 16;
      *reinterpret_cast<float4*>(stage + 2 * TG_TILE + off) = make_float4(w0[0], w0[1], w0[2], w0[3]);
      if (DUAL) *reinterpret_cast<float4*>(stage + 3 * TG_TILE + off) = make_float4(w1[0], w1[1], w1[2], w1[3]);
    }
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      float a0[4], a1[4];
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const int j = 4 * g + jj;
        const float v = va[j];
        a0[jj] = oka[j] ? to_tf32(v) : 0.f;
        if (DUAL) {
          float w = two_bases ? va1[j] : v;
          if (p.trA == TRA_SQUARE) w = w * w;
          else if (p.trA == TRA_SIGN) w *= va1[j];
          a1[jj] = oka[j] ? to_tf32(w) : 0.f;
        }
      }
      const int off = (kq * 4 + g) * TG_CHUNK + r * 16;
      *reinterpret_cast<float4*>(stage + off) = make_float4(a0[0], a0[1], a0[2], a0[3]);
      if (DUAL) *reinterpret_cast<float4*>(stage + TG_TILE + off) = make_float4(a1[0], a1[1], a1[2], a1[3]);
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    if (c.warp == 0) {
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sa = c.sbase + st * TG_STAGE;
        const uint32_t idesc = umma_idesc_tf32(NP);
#pragma unroll
        for (int ks = 0; ks < TG_K / 8; ++ks) {
          umma_tf32(c.tmem, umma_desc(sa + 2 * ks * TG_CHUNK, TG_CHUNK, 128), umma_desc(sa + 2 * TG_TILE + 2 * ks * TG_CHUNK, TG_CHUNK, 128),
                    idesc, (ch | ks) != 0);
          if (DUAL)
            umma_tf32(c.tmem + 128, umma_desc(sa + TG_TILE + 2 * ks * TG_CHUNK, TG_CHUNK, 128),
                      umma_desc(sa + 3 * TG_TILE + 2 * ks * TG_CHUNK, TG_CHUNK, 128), idesc, (ch | ks) != 0);
        }
        umma_commit(c.bars + 8 * st);
      }
      __syncwarp();
    }
  }
  // ---- accumulators -> shared memory ([n][m], the ring is dead once the last commit has arrived) -> rolled epilogue
  float* Cs0 = reinterpret_cast<float*>(smem);
  float* Cs1 = reinterpret_cast<float*>(smem + TG_STAGE);
  if (nchunk > 0) {
    if (nchunk >= 2) mbar_wait(c.bars + 8 * ((nchunk - 2) & 1), ((nchunk - 2) >> 1) & 1, status, 21);
    mbar_wait(c.bars + 8 * ((nchunk - 1) & 1), ((nchunk - 1) >> 1) & 1, status, 21);
    tc_fence_after();
    const int q = c.warp & 3, half = c.warp >> 2, row = q * 32 + c.lane;
    const uint32_t la = c.tmem + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
    for (int g = 0; g < 4; ++g) {
      const int col0 = half * 64 + g * 16;
      if (col0 >= NP) break;
      float v0[16], v1[16];
      tmem_ld16(la + col0, v0);
      if (DUAL) tmem_ld16(la + 128 + col0, v1);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        Cs0[(col0 + j) * 128 + row] = v0[j];
        if (DUAL) Cs1[(col0 + j) * 128 + row] = v1[j];
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  // 4-way unrolled: the read-modify-write / noise loads of four outputs overlap (a fully rolled loop pays one global
  // round trip per output, a fully unrolled one is instruction-fetch bound: Philox / log / sincos per element)
#pragma unroll 4
  for (int e = tid; e < 128 * ncol; e += 256) {
    const int mi = e & 127, ni = e >> 7;
    const int m = m0 + mi, n = n0 + ni;
    if (m >= Mtot) continue;
    const float a0 = nchunk > 0 ? Cs0[e] : 0.f, a1 = (DUAL && nchunk > 0) ? Cs1[e] : 0.f;
    if (split) {
      if (nchunk > 0) {
        atomicAdd(p.part + (long long)m * p.N + n, a0);
        if (DUAL) atomicAdd(p.part + (long long)Mtot * p.N + (long long)m * p.N + n, a1);
      }
    } else {
      gemm_epilogue<EPI>(p, s, m, n, a0, a1);
    }
  }
  tg_teardown(c, 256);
}

template <bool DUAL, int EPI>
__global__ void tc_splitk_epilogue_kernel(const ConvGemm p) {
  const long long total = (long long)p.B * p.P * p.N;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int m = (int)(i / p.N), n = (int)(i - (long long)m * p.N);
    gemm_epilogue<EPI>(p, 0, m, n, p.part[i], DUAL ? p.part[total + i] : 0.f);
  }
}

static int* g_tg_status = nullptr;  // device word: first mbarrier time-out code of a TF32 kernel (0 = ok)

static int* tg_status_word() {
  if (!g_tg_status) {
    cudaMalloc(&g_tg_status, sizeof(int));
    cudaMemset(g_tg_status, 0, sizeof(int));
  }
  return g_tg_status;
}
int tc_gemm_status() {
  int v = 0;
  if (g_tg_status) cudaMemcpy(&v, g_tg_status, sizeof(int), cudaMemcpyDeviceToHost);
  return v;
}

template <bool DUAL, int EPI>
static void tg_launch(const ConvGemm& p, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(tc_gemm_kernel<DUAL, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, TG_SMEM);
    configured = true;
  }
  const int Mtot = p.B * p.P;
  const int z = p.ksplit > 1 ? p.ksplit : p.S;
  ++g_launch_count;
  tc_gemm_kernel<DUAL, EPI><<<dim3((Mtot + 127) / 128, (p.N + 127) / 128, z), 256, TG_SMEM, st>>>(p, tg_status_word());
  if (p.ksplit > 1) {
    const long long total = (long long)Mtot * p.N;
    ++g_launch_count;
    tc_splitk_epilogue_kernel<DUAL, EPI><<<(unsigned)std::min<long long>((total + 255) / 256, 148 * 8), 256, 0, st>>>(p);
  }
}

// K split for single-sample launches whose grid would leave most SMs idle (the 2400 -> 64 Linear at training batches)
static int tg_ksplit(const ConvGemm& p) {
  if (p.S != 1 || p.part == nullptr) return 1;
  const int ctas = ((p.B * p.P + 127) / 128) * ((p.N + 127) / 128);
  if (ctas >= 32 || p.K < 8 * TG_K) return 1;
  // the scratch tile is sized for the SIMT path: 2 x 74 x 128 x 32 floats
  if (2ll * p.B * p.P * p.N > SPLITK_SCRATCH_FLOATS) return 1;
  return std::max(1, std::min(p.K / (2 * TG_K), 148 / ctas));
}

void launch_conv_gemm_tc(const ConvGemm& p0, int epi, cudaStream_t st) {
  ConvGemm p = p0;
  p.ksplit = tg_ksplit(p);
  if (p.ksplit > 1) {
    const bool dual = epi != EPI_FWD_PLAIN && epi != EPI_DX_PLAIN;
    cudaMemsetAsync(p.part, 0, sizeof(float) * (size_t)p.B * p.P * p.N * (dual ? 2 : 1), st);
  }
  switch (epi) {
    case EPI_FWD_PLAIN: tg_launch<false, EPI_FWD_PLAIN>(p, st); break;
    case EPI_FWD_LRT: tg_launch<true, EPI_FWD_LRT>(p, st); break;
    case EPI_FWD_FLIPOUT: tg_launch<true, EPI_FWD_FLIPOUT>(p, st); break;
    case EPI_DX_PLAIN: tg_launch<false, EPI_DX_PLAIN>(p, st); break;
    case EPI_DX_LRT: tg_launch<true, EPI_DX_LRT>(p, st); break;
    case EPI_DX_FLIPOUT: tg_launch<true, EPI_DX_FLIPOUT>(p, st); break;
  }
}

// ------------------------------------------------------------------------------------------------
// weight gradient: D[k][co] = sum_m tr(A[m][k]) * G[m][co];  MMA M = k tile (128), N = co tile, K = rows m
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 1) tc_dw_kernel(const ConvDw p, int rows_per_split, int* status) {
  extern __shared__ __align__(128) unsigned char smem[];
  const TgCtx c = tg_setup(smem, 128);
  const int tid = threadIdx.x;
  const int kt0 = blockIdx.x * 128, co0 = blockIdx.y * 128;
  const int Mtot = p.B * p.P;
  const int mbeg = blockIdx.z * rows_per_split, mend = min(Mtot, mbeg + rows_per_split);
  const int ncol = min(128, p.N - co0), NP = (ncol + 15) & ~15;
  const int nchunk = mend > mbeg ? (mend - mbeg + TG_K - 1) / TG_K : 0;

  // thread = (k column / co column r, half mq of the stage's 32 rows): the column constants never change
  const int r = tid & 127, mq = tid >> 7;
  const int k = kt0 + r;
  const int kmode = k < p.K ? 0 : (k == p.K ? 1 : 2);  // 0 gather, 1 bias column (A = 1), 2 out of range
  const int kc = min(k, p.K - 1);
  const int dhw = __ldg(p.a.kdhw + kc);
  const int ko = __ldg(p.a.koff + kc);
  const int kdh = (int)(short)(dhw & 0xffff), kdw = dhw >> 16;
  const int kcc = p.trA == TRA_SIGN ? __ldg(p.a.kci + kc) : 0;
  const int co = co0 + r;
  const bool cov = r < ncol;

  for (int ch = 0; ch < nchunk; ++ch) {
    const int st = ch & 1;
    unsigned char* stage = smem + st * TG_STAGE;
    if (ch >= 2) {
      mbar_wait(c.bars + 8 * st, ((ch >> 1) - 1) & 1, status, 22);
      tc_fence_after();
    }
    const int mb = mbeg + ch * TG_K + mq * 16;
    int b = mb / p.P, pp = mb - b * p.P;
    // all loads of the stage first (latency-bound), then transforms + TF32 rounding + stores
    float va[16], vs[16], vg[16];
    bool oka[16], okg[16], rvv[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int m = mb + j;
      const bool rv = m < mend;
      const int oh = pp / p.Wrow, ow = pp - oh * p.Wrow;
      oka[j] = rv && kmode == 0 && (unsigned)(oh + kdh) < (unsigned)p.a.Hin && (unsigned)(ow + kdw) < (unsigned)p.a.Win;
      rvv[j] = rv;
      const long long rowbase = (long long)b * p.a.img_stride + (long long)oh * p.a.sH + (long long)ow * p.a.sW;
      va[j] = __ldg(p.a.base0 + (oka[j] ? rowbase + ko : 0));
      if (p.trA == TRA_SIGN) vs[j] = __ldg(p.sign_in + (oka[j] ? (long long)b * p.sign_C + kcc : 0));
      okg[j] = rv && cov;
      vg[j] = __ldg(p.G + (okg[j] ? ((long long)b * p.N + co) * p.P + pp : 0));
      if (++pp == p.P) { pp = 0; ++b; }
    }
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      float av[4], gv[4];
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const int j = 4 * g + jj;
        float v = va[j];
        if (p.trA == TRA_SQUARE) v = v * v;
        else if (p.trA == TRA_SIGN) v *= vs[j];
        av[jj] = oka[j] ? to_tf32(v) : ((rvv[j] && kmode == 1) ? 1.0f : 0.f);
        gv[jj] = okg[j] ? to_tf32(vg[j]) : 0.f;
      }
      const int off = (mq * 4 + g) * TG_CHUNK + r * 16;
      *reinterpret_cast<float4*>(stage + off) = make_float4(av[0], av[1], av[2], av[3]);
      *reinterpret_cast<float4*>(stage + 2 * TG_TILE + off) = make_float4(gv[0], gv[1], gv[2], gv[3]);
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    if (c.warp == 0) {
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sa = c.sbase + st * TG_STAGE;
        const uint32_t idesc = umma_idesc_tf32(NP);
#pragma unroll
        for (int ks = 0; ks < TG_K / 8; ++ks)
          umma_tf32(c.tmem, umma_desc(sa + 2 * ks * TG_CHUNK, TG_CHUNK, 128), umma_desc(sa + 2 * TG_TILE + 2 * ks * TG_CHUNK, TG_CHUNK, 128),
                    idesc, (ch | ks) != 0);
        umma_commit(c.bars + 8 * st);
      }
      __syncwarp();
    }
  }
  if (nchunk > 0) {
    if (nchunk >= 2) mbar_wait(c.bars + 8 * ((nchunk - 2) & 1), ((nchunk - 2) >> 1) & 1, status, 23);
    mbar_wait(c.bars + 8 * ((nchunk - 1) & 1), ((nchunk - 1) >> 1) & 1, status, 23);
    tc_fence_after();
    // thread = k row (TMEM lane) x column half: fp32 atomics into the flat gradient buffer (lanes -> consecutive k)
    const int q = c.warp & 3, half = c.warp >> 2;
    const int kr = kt0 + q * 32 + c.lane;
    const uint32_t la = c.tmem + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
    for (int g = 0; g < 4; ++g) {
      const int col0 = half * 64 + g * 16;
      if (col0 >= NP) break;
      float v[16];
      tmem_ld16(la + col0, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int cc = co0 + col0 + j;
        if (cc < p.N) {
          if (kr < p.K) {
            atomicAdd(p.gw + (long long)cc * p.K + kr, v[j]);
          } else if (kr == p.K) {
            if (p.gb) atomicAdd(p.gb + cc, v[j]);
            if (p.gb2) atomicAdd(p.gb2 + cc, v[j]);
          }
        }
      }
    }
  }
  tg_teardown(c, 128);
}

void launch_conv_dw_tc(const ConvDw& p, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(tc_dw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TG_SMEM);
    configured = true;
  }
  const int Mtot = p.B * p.P;
  const int gx = (p.K + 1 + 127) / 128, gy = (p.N + 127) / 128;
  int split = std::max(1, std::min((Mtot + TG_K - 1) / TG_K, (148 + gx * gy - 1) / (gx * gy)));
  int rows = (Mtot + split - 1) / split;
  rows = (rows + TG_K - 1) / TG_K * TG_K;
  split = (Mtot + rows - 1) / rows;
  ++g_launch_count;
  tc_dw_kernel<<<dim3(gx, gy, split), 256, TG_SMEM, st>>>(p, rows, tg_status_word());
}

}  // namespace brl
