// fp32 SIMT implicit-GEMM kernels of the parity engine: forward / input-gradient (conv_gemm_kernel) and
// weight-gradient (conv_dw_kernel) passes of every conv / linear layer, with the variational rules fused in
// the gather prologue and the epilogue.
//
// Loads are branch-free and software-pipelined: the gather tables of a K-tile are read unconditionally, the
// data loads use a clamped (always valid) address + select, and the next K-tile is fetched into registers
// while the current one is multiplied out of shared memory.  Small-M / long-K layers (the 2400->64 Linear at
// training batch sizes) are split over K (blockIdx.z) with fp32 atomics into a scratch tile followed by a
// one-thread-per-output epilogue kernel.
#include <algorithm>
#include <atomic>
#include <cstdlib>

#include "brl_kernels.cuh"
#include "brl_philox.cuh"
#include "brl_gemm_epi.cuh"

namespace brl {

extern std::atomic<long long> g_launch_count;

constexpr int BN = 32, BK = 16;

// Tile shapes (BM rows x 32 columns, TM x TN outputs per thread):
//   128 x 32, 4 x 4, 256 threads  large row counts (predictive batches): FFMA-bound, wide micro-tiles
//    32 x 32, 2 x 2, 256 threads  training batches: a 256-window layer is a few hundred tiles, i.e. one or two warps per
//                                 scheduler with 4 x 4 micro-tiles, and the kernel then runs at the latency of its own
//                                 dependent chain (ncu: 22 % issue slots, 3.7 long-scoreboard stalls per issue); quartering
//                                 the per-thread tile gives 4x the warps and a third of the serial instructions per warp
template <bool DUAL, int EPI, int BM, int TM, int TN, int BKT>
__global__ void __launch_bounds__((BM / TM) * (BN / TN)) conv_gemm_kernel(const ConvGemm p) {
  constexpr int NT = (BM / TM) * (BN / TN);  // threads
  constexpr int AKS = NT / BM;        // k-rows of the A tile staged per pass
  constexpr int ALD = BKT / AKS;       // A loads per thread
  constexpr int BROWS = NT / BN;      // k-rows of the B tile staged per pass
  constexpr int BLD = BKT / BROWS;     // B loads per thread
  constexpr int TXN = BN / TN;        // threads along n
  static_assert(ALD >= 1 && BLD >= 1 && (TM == 4 || TM == 2) && (TN == 4 || TN == 2), "tile shape");
  // K-loop tiles and (afterwards) the accumulator tile of the rolled epilogue share one buffer
  constexpr int LOOP_FLOATS = (DUAL ? 2 : 1) * (BKT * BM + BKT * BN);
  constexpr int EPI_FLOATS = (DUAL ? 2 : 1) * BM * BN;
  __shared__ __align__(16) float sm[LOOP_FLOATS > EPI_FLOATS ? LOOP_FLOATS : EPI_FLOATS];
  float(*As0)[BM] = reinterpret_cast<float(*)[BM]>(sm);
  float(*Bs0)[BN] = reinterpret_cast<float(*)[BN]>(sm + BKT * BM);
  float(*As1)[BM] = reinterpret_cast<float(*)[BM]>(sm + BKT * BM + BKT * BN);
  float(*Bs1)[BN] = reinterpret_cast<float(*)[BN]>(sm + 2 * BKT * BM + BKT * BN);

  const bool split = p.ksplit > 1;
  const int s = split ? 0 : blockIdx.z;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int tid = threadIdx.x;
  const int Mtot = p.B * p.P;
  int kbeg = 0, kend = p.K;
  if (split) {
    const int per = ((p.K + p.ksplit - 1) / p.ksplit + BKT - 1) / BKT * BKT;
    kbeg = blockIdx.z * per;
    kend = min(p.K, kbeg + per);
  }

  const int ar = tid & (BM - 1), ak0 = tid / BM;
  const int am = m0 + ar;
  const bool arv = am < Mtot;
  const int ab = arv ? am / p.P : 0;
  const int app = arv ? am - ab * p.P : 0;
  const int aoh = app / p.Wrow, aow = app - aoh * p.Wrow;
  const long long aimg = p.a.per_sample ? (long long)s * p.B + ab : ab;
  const long long rowbase = aimg * p.a.img_stride + (long long)aoh * p.a.sH + (long long)aow * p.a.sW;
  const float* sgn = p.sign_in + ((long long)s * p.B + ab) * p.sign_C;
  const int bn = tid & (BN - 1), bk0 = tid >> 5;
  const int bnn = n0 + bn;
  const bool bnv = bnn < p.N;
  const long long bcol = (long long)(bnv ? bnn : 0) * p.nB;
  const float* W0 = p.W0 + (long long)s * p.ws0;
  const float* W1 = DUAL ? p.W1 + (long long)s * p.ws1 : nullptr;
  const int tx = tid % TXN, ty = tid / TXN;

  float ra0[ALD], ra1[ALD], rb0[BLD], rb1[BLD];
  auto fetch = [&](int k0) {
    int ko[ALD], kc[ALD];
    bool ok[ALD];
#pragma unroll
    for (int j = 0; j < ALD; ++j) {  // gather tables: unconditional, clamped
      const int k = k0 + ak0 + AKS * j;
      kc[j] = min(k, p.K - 1);
      const int dhw = __ldg(p.a.kdhw + kc[j]);
      ko[j] = __ldg(p.a.koff + kc[j]);
      const int ih = aoh + (int)(short)(dhw & 0xffff), iw = aow + (dhw >> 16);
      ok[j] = arv && k < kend && (unsigned)ih < (unsigned)p.a.Hin && (unsigned)iw < (unsigned)p.a.Win;
    }
#pragma unroll
    for (int j = 0; j < ALD; ++j) {  // data: always-valid address + select (no branches between loads)
      const long long off = ok[j] ? rowbase + ko[j] : 0;
      const float v = __ldg(p.a.base0 + off);
      ra0[j] = ok[j] ? v : 0.f;
      if (DUAL) {
        float w = (p.a.base1 == p.a.base0) ? v : __ldg(p.a.base1 + off);
        if (p.trA == TRA_SQUARE) w = w * w;
        else if (p.trA == TRA_SIGN) w *= __ldg(sgn + (ok[j] ? __ldg(p.a.kci + kc[j]) : 0));
        ra1[j] = ok[j] ? w : 0.f;
      }
    }
#pragma unroll
    for (int j = 0; j < BLD; ++j) {
      const int k = k0 + bk0 + BROWS * j;
      const int kcl = min(k, p.K - 1);
      const bool okb = bnv && k < kend;
      const long long off = (p.kB ? (long long)__ldg(p.kB + kcl) : (long long)kcl) + bcol;
      const float v = __ldg(W0 + off);
      rb0[j] = okb ? v : 0.f;
      if (DUAL) {
        float w = __ldg(W1 + off);
        if (p.trB == TRB_SQUARE) w = w * w;
        else if (p.trB == TRB_MINUS_W0) w -= v;
        rb1[j] = okb ? w : 0.f;
      }
    }
  };

  float acc0[TM][TN], acc1[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc0[i][j] = acc1[i][j] = 0.f;
  auto frag = [&](const float* src, float (&dst)[4], int n) {  // n = 2 or 4 consecutive floats
    if (n == 4) {
      const float4 v = *reinterpret_cast<const float4*>(src);
      dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
    } else {
      const float2 v = *reinterpret_cast<const float2*>(src);
      dst[0] = v.x; dst[1] = v.y;
    }
  };

  if (kbeg < kend) fetch(kbeg);
  for (int k0 = kbeg; k0 < kend; k0 += BKT) {
#pragma unroll
    for (int j = 0; j < ALD; ++j) {
      As0[ak0 + AKS * j][ar] = ra0[j];
      if (DUAL) As1[ak0 + AKS * j][ar] = ra1[j];
    }
#pragma unroll
    for (int j = 0; j < BLD; ++j) {
      Bs0[bk0 + BROWS * j][bn] = rb0[j];
      if (DUAL) Bs1[bk0 + BROWS * j][bn] = rb1[j];
    }
    __syncthreads();
    if (k0 + BKT < kend) fetch(k0 + BKT);  // next tile's loads fly while this one is multiplied
#pragma unroll
    for (int kk = 0; kk < BKT; ++kk) {
      float av[4], bv[4];
      frag(&As0[kk][ty * TM], av, TM);
      frag(&Bs0[kk][tx * TN], bv, TN);
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc0[i][j] = fmaf(av[i], bv[j], acc0[i][j]);
      if (DUAL) {
        float av1[4], bv1[4];
        frag(&As1[kk][ty * TM], av1, TM);
        frag(&Bs1[kk][tx * TN], bv1, TN);
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) acc1[i][j] = fmaf(av1[i], bv1[j], acc1[i][j]);
      }
    }
    __syncthreads();
  }

  // ---- rolled epilogue: park the accumulators in shared memory ([n][m], conflict-free) and let every thread
  // walk 16 outputs in a loop.  (Unrolling the Philox / log / sincos epilogue 16x made the kernel 270 KB of
  // SASS and instruction-fetch bound.)  Consecutive threads own consecutive rows -> coalesced stores.
  float* Cs0 = sm;
  float* Cs1 = sm + BM * BN;
#pragma unroll
  for (int j = 0; j < TN; ++j)
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      Cs0[(tx * TN + j) * BM + ty * TM + i] = acc0[i][j];
      if (DUAL) Cs1[(tx * TN + j) * BM + ty * TM + i] = acc1[i][j];
    }
  __syncthreads();
#pragma unroll 1
  for (int e = tid; e < BM * BN; e += NT) {
    const int mi = e & (BM - 1), ni = e / BM;
    const int m = m0 + mi, n = n0 + ni;
    if (m >= Mtot || n >= p.N) continue;
    const float a0 = Cs0[e], a1 = DUAL ? Cs1[e] : 0.f;
    if (split) {
      atomicAdd(p.part + (long long)m * p.N + n, a0);
      if (DUAL) atomicAdd(p.part + (long long)Mtot * p.N + (long long)m * p.N + n, a1);
    } else {
      gemm_epilogue<EPI>(p, s, m, n, a0, a1);
    }
  }
}

template <bool DUAL, int EPI>
__global__ void splitk_epilogue_kernel(const ConvGemm p) {
  const long long total = (long long)p.B * p.P * p.N;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int m = (int)(i / p.N), n = (int)(i - (long long)m * p.N);
    gemm_epilogue<EPI>(p, 0, m, n, p.part[i], DUAL ? p.part[total + i] : 0.f);
  }
}

static bool small_tiles(const ConvGemm& p) {
  const long long ctas128 = (long long)((p.B * p.P + 127) / 128) * ((p.N + BN - 1) / BN) * (p.ksplit > 1 ? p.ksplit : p.S);
  return ctas128 < 4 * 148;
}

template <bool DUAL, int EPI>
static void launch_one(const ConvGemm& p, cudaStream_t st) {
  const int Mtot = p.B * p.P;
  const int z = p.ksplit > 1 ? p.ksplit : p.S;
  ++g_launch_count;
  static const int small_shape = getenv("BRL_SMALL_TILE") ? atoi(getenv("BRL_SMALL_TILE")) : 42;  // experiment knob
  if (small_tiles(p)) {
    const dim3 grid((Mtot + 31) / 32, (p.N + BN - 1) / BN, z);
    if (small_shape == 42) conv_gemm_kernel<DUAL, EPI, 32, 4, 2, 16><<<grid, 128, 0, st>>>(p);
    else if (small_shape == 423) conv_gemm_kernel<DUAL, EPI, 32, 4, 2, 32><<<grid, 128, 0, st>>>(p);
    else if (small_shape == 223) conv_gemm_kernel<DUAL, EPI, 32, 2, 2, 32><<<grid, 256, 0, st>>>(p);
    else conv_gemm_kernel<DUAL, EPI, 32, 2, 2, 16><<<grid, 256, 0, st>>>(p);
  } else
    conv_gemm_kernel<DUAL, EPI, 128, 4, 4, 16><<<dim3((Mtot + 127) / 128, (p.N + BN - 1) / BN, z), 256, 0, st>>>(p);
  if (p.ksplit > 1) {
    const long long total = (long long)Mtot * p.N;
    ++g_launch_count;
    splitk_epilogue_kernel<DUAL, EPI><<<(unsigned)std::min<long long>((total + 255) / 256, 148 * 8), 256, 0, st>>>(p);
  }
}

// picks a K split for single-sample launches whose grid would leave most SMs idle
int conv_gemm_ksplit(const ConvGemm& p) {
  if (p.S != 1 || p.part == nullptr) return 1;
  const int ctas = ((p.B * p.P + 127) / 128) * ((p.N + BN - 1) / BN);  // scratch is sized for < 74 tiles of 128x32
  if (ctas >= 74 || p.K < 8 * BK) return 1;
  static const int ks_min32 = getenv("BRL_KSPLIT_MIN32") ? atoi(getenv("BRL_KSPLIT_MIN32")) : 0;  // experiment knob
  if (ks_min32 > 0 && ((p.B * p.P + 31) / 32) * ((p.N + BN - 1) / BN) >= ks_min32) return 1;
  return std::max(1, std::min(p.K / (4 * BK), 296 / ctas));
}

// the epilogue alone, over partial sums some other kernel has accumulated in p.part ([2][B * P][N]): the fc layer of the level-fused
// training kernels (brl_tc_train.cu) contracts on the tensor pipe and shares bias / noise / activation handling with this path
void launch_splitk_epilogue(const ConvGemm& p, int epi, cudaStream_t st) {
  const long long total = (long long)p.B * p.P * p.N;
  const unsigned grid = (unsigned)std::min<long long>((total + 255) / 256, 148 * 8);
  ++g_launch_count;
  switch (epi) {
    case EPI_FWD_PLAIN: splitk_epilogue_kernel<false, EPI_FWD_PLAIN><<<grid, 256, 0, st>>>(p); break;
    case EPI_FWD_LRT: splitk_epilogue_kernel<true, EPI_FWD_LRT><<<grid, 256, 0, st>>>(p); break;
    case EPI_FWD_FLIPOUT: splitk_epilogue_kernel<true, EPI_FWD_FLIPOUT><<<grid, 256, 0, st>>>(p); break;
    default: break;
  }
}

void launch_conv_gemm(const ConvGemm& p0, int epi, cudaStream_t st) {
  ConvGemm p = p0;
  p.ksplit = conv_gemm_ksplit(p);
  if (p.ksplit > 1) {
    const bool dual = epi != EPI_FWD_PLAIN && epi != EPI_DX_PLAIN;
    cudaMemsetAsync(p.part, 0, sizeof(float) * (size_t)p.B * p.P * p.N * (dual ? 2 : 1), st);
  }
  switch (epi) {
    case EPI_FWD_PLAIN: launch_one<false, EPI_FWD_PLAIN>(p, st); break;
    case EPI_FWD_LRT: launch_one<true, EPI_FWD_LRT>(p, st); break;
    case EPI_FWD_FLIPOUT: launch_one<true, EPI_FWD_FLIPOUT>(p, st); break;
    case EPI_DX_PLAIN: launch_one<false, EPI_DX_PLAIN>(p, st); break;
    case EPI_DX_LRT: launch_one<true, EPI_DX_LRT>(p, st); break;
    case EPI_DX_FLIPOUT: launch_one<true, EPI_DX_FLIPOUT>(p, st); break;
  }
}

// ------------------------------------------------------------------------------------------------
// weight gradient: C[co][k] += sum_m G[m][co] * tr(A[m][k]); split over row ranges, fp32 atomics
// ------------------------------------------------------------------------------------------------
constexpr int DW_CO = 32, DW_K = 128, DW_M = 16, DW_PAD = 4;

template <bool DUAL>
__global__ void __launch_bounds__(256, 2) conv_dw_kernel(const ConvDw p, int rows_per_split) {
  __shared__ __align__(16) float Gs[DUAL ? 2 : 1][DW_M][DW_CO];
  __shared__ __align__(16) float As[DUAL ? 2 : 1][DW_M][DW_K + DW_PAD];
  const int tid = threadIdx.x;
  const int kt0 = blockIdx.x * DW_K, co0 = blockIdx.y * DW_CO;
  const int Mtot = p.B * p.P;
  const int mbeg = blockIdx.z * rows_per_split;
  const int mend = min(Mtot, mbeg + rows_per_split);
  const int r = tid & 15, c0 = tid >> 4;
  const int tx = tid & 31, ty = tid >> 5;
  const int tr1 = DUAL ? p.trA1 : p.trA;  // the transform that needs the sign tensor, if any

  // per-thread column constants (the 8 k-columns and 2 co-columns this thread stages never change)
  int ko[8], kdh[8], kdw[8], kcc[8], kmode[8];  // kmode: 0 gather, 1 bias column (A = 1), 2 out of range
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int k = kt0 + c0 + 16 * j;
    kmode[j] = k < p.K ? 0 : (k == p.K ? 1 : 2);
    const int kc = min(k, p.K - 1);
    const int dhw = __ldg(p.a.kdhw + kc);
    ko[j] = __ldg(p.a.koff + kc);
    kdh[j] = (int)(short)(dhw & 0xffff);
    kdw[j] = dhw >> 16;
    kcc[j] = tr1 == TRA_SIGN ? __ldg(p.a.kci + kc) : 0;
  }
  float rg[2], rg1[2], ra[8], ra1[8];
  auto fetch = [&](int mb) {
    const int m = mb + r;
    const bool rv = m < mend;
    const int b = rv ? m / p.P : 0;
    const int pp = rv ? m - b * p.P : 0;
    const int oh = pp / p.Wrow, ow = pp - oh * p.Wrow;
    const long long rowbase = (long long)b * p.a.img_stride + (long long)oh * p.a.sH + (long long)ow * p.a.sW;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int co = co0 + c0 + 16 * j;
      const bool ok = rv && co < p.N;
      const long long gi = ok ? ((long long)b * p.N + co) * p.P + pp : 0;
      const float v = __ldg(p.G + gi);
      rg[j] = ok ? v : 0.f;
      if (DUAL) {
        const float v1 = __ldg(p.G1 + gi);
        rg1[j] = ok ? v1 : 0.f;
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const bool ok = rv && kmode[j] == 0 && (unsigned)(oh + kdh[j]) < (unsigned)p.a.Hin &&
                      (unsigned)(ow + kdw[j]) < (unsigned)p.a.Win;
      const float v = __ldg(p.a.base0 + (ok ? rowbase + ko[j] : 0));
      float t = v;
      if (tr1 == TRA_SQUARE) t = v * v;
      else if (tr1 == TRA_SIGN) t = v * __ldg(p.sign_in + (ok ? (long long)b * p.sign_C + kcc[j] : 0));
      const float one = (rv && kmode[j] == 1) ? 1.0f : 0.f;
      if (DUAL) {
        ra[j] = ok ? v : one;
        ra1[j] = ok ? t : one;
      } else {
        ra[j] = ok ? t : one;
      }
    }
  };

  float acc[4][4], acc1[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = acc1[i][j] = 0.f;

  if (mbeg < mend) fetch(mbeg);
  for (int mb = mbeg; mb < mend; mb += DW_M) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      Gs[0][r][c0 + 16 * j] = rg[j];
      if (DUAL) Gs[DUAL ? 1 : 0][r][c0 + 16 * j] = rg1[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      As[0][r][c0 + 16 * j] = ra[j];
      if (DUAL) As[DUAL ? 1 : 0][r][c0 + 16 * j] = ra1[j];
    }
    __syncthreads();
    if (mb + DW_M < mend) fetch(mb + DW_M);
#pragma unroll
    for (int mm = 0; mm < DW_M; ++mm) {
      const float4 g = *reinterpret_cast<const float4*>(&Gs[0][mm][ty * 4]);
      const float4 a = *reinterpret_cast<const float4*>(&As[0][mm][tx * 4]);
      const float gv[4] = {g.x, g.y, g.z, g.w}, av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(gv[i], av[j], acc[i][j]);
      if (DUAL) {
        const float4 g1 = *reinterpret_cast<const float4*>(&Gs[DUAL ? 1 : 0][mm][ty * 4]);
        const float4 a1 = *reinterpret_cast<const float4*>(&As[DUAL ? 1 : 0][mm][tx * 4]);
        const float gv1[4] = {g1.x, g1.y, g1.z, g1.w}, av1[4] = {a1.x, a1.y, a1.z, a1.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc1[i][j] = fmaf(gv1[i], av1[j], acc1[i][j]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int co = co0 + ty * 4 + i;
    if (co >= p.N) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = kt0 + tx * 4 + j;
      if (k < p.K) {
        atomicAdd(p.gw + (long long)co * p.K + k, acc[i][j]);
        if (DUAL) atomicAdd(p.gw1 + (long long)co * p.K + k, acc1[i][j]);
      } else if (k == p.K) {
        if (p.gb) atomicAdd(p.gb + co, acc[i][j]);
        if (p.gb2) atomicAdd(p.gb2 + co, acc[i][j]);
        if (DUAL && p.gb1) atomicAdd(p.gb1 + co, acc1[i][j]);
      }
    }
  }
}

void launch_conv_dw(const ConvDw& p, cudaStream_t st) {
  const int Mtot = p.B * p.P;
  const int gx = (p.K + 1 + DW_K - 1) / DW_K, gy = (p.N + DW_CO - 1) / DW_CO;
  static const int dw_ctas = getenv("BRL_DW_CTAS") ? atoi(getenv("BRL_DW_CTAS")) : 148;  // experiment knob
  int split = std::max(1, std::min((Mtot + 31) / 32, (dw_ctas + gx * gy - 1) / (gx * gy)));
  int rows = (Mtot + split - 1) / split;
  rows = (rows + DW_M - 1) / DW_M * DW_M;
  split = (Mtot + rows - 1) / rows;
  ++g_launch_count;
  if (p.G1) conv_dw_kernel<true><<<dim3(gx, gy, split), 256, 0, st>>>(p, rows);
  else conv_dw_kernel<false><<<dim3(gx, gy, split), 256, 0, st>>>(p, rows);
}

}  // namespace brl
