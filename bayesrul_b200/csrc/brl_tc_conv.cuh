// tcgen05 conv-stack kernel (included by brl_tc.cu after the PTX wrappers and geometry constants).
//
// One persistent CTA per SM, 512 threads = two independent 256-thread groups.  Each group owns one 128-row tile
// (4 windows x 32 rows; a warp = one window) at a time, its own 66 KB activation region, 256 TMEM columns, three
// mbarriers and one named barrier, so while one group waits for its MMAs the other runs its epilogue.  The fp16
// weights of the current MC sample (86 KB) are shared by both groups and stay resident.
//
// Activation region of a group (K-major, no swizzle: [8-channel chunk][132 rows][16 B]):
//   M1  16 chunks  module-1 output, 4 branches x 32 channels (27 real + const-1 channel + zeros)
//   M1P 16 chunks  MaxPool1d(3,1,1)(M1)  -- produced in the phase-A epilogue with warp shuffles (row = lane)
//   T2 / T3 (8 + 8 chunks) alias M1   (written after the MMAs reading M1 have completed)
//   X / XP  (4 + 4 chunks) alias M1P  (read by phase A only; M1P is written after phase A completed)
// MMA count per tile is kept low because every M=128 MMA re-reads its 4 KB A slice from shared memory whatever N is:
//   phase A  convs that share an input shift share one MMA (N = 96 / 64 / 32 for |shift| = 0 / 1 / 2)      -> 16 MMAs
//   phase B  [b1 | b2a | b3a] as one N = 144 GEMM over M1, b4 (N = 32) over pooled M1                          -> 16 MMAs
//   phase C  all taps of a k3 / k5 conv are concatenated along N (N = 48 / 80, un-shifted A) and the tap shift is
//            applied afterwards in registers: out[t] = sum_tap P_tap[t + tap - pad] via __shfl_up / __shfl_down   ->  8 MMAs
// Biases ride in the MMAs: input feature 18 is a constant 1 whose weight row (centre tap only) holds the bias;
// conv1's column 27 reproduces the constant into M1 (and, pooled, into M1P) for the module-2 1x1 convs.
// The next tile's windows are prefetched into registers while the current tile is in flight (raw fp16 staging buffer).
#pragma once

struct ConvArgs {
  const float* x;             // [B,30,18]
  const unsigned char* blob;  // [S or 1][BLOB_BYTES]
  long long blob_stride;      // BLOB_BYTES or 0
  unsigned char* feat;        // [S][NT128][300][128][16 B]
  int B, S, ntile4, ntile128;
  float keep4;                // dropout keep of the branch sites (1 = off)
  NoiseRef drop[12];
  int* status;
};

__device__ __forceinline__ void group_sync(int grp) { asm volatile("bar.sync %0, 256;" ::"r"(grp + 1) : "memory"); }

// ReLU (+bias) (+dropout) on 16 fp32 accumulator columns
template <bool DROP, bool BIAS>
__device__ __forceinline__ void act16(float (&v)[16], const float* bias, bool live, const ConvArgs& a, int layer, int s,
                                      int gw, int t, int ch0, int nvalid) {
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    float u = fmaxf(BIAS ? v[j] + bias[j] : v[j], 0.f);
    if (DROP) {
      if (ch0 + j < nvalid && live) {
        const NoiseRef& nz = a.drop[layer];
        const int e = (ch0 + j) * 30 + t;
        const bool keep = nz.ptr ? nz.ptr[((long long)s * a.B + gw) * (nvalid * 30) + e] != 0.f
                                 : philox_uniform(nz.seed, nz.kind, nz.site, nz.sample0 + s, nz.window0 + gw, e) < a.keep4;
        u = keep ? u / a.keep4 : 0.f;
      }
    }
    v[j] = u;
  }
}
__device__ __forceinline__ void pack16(const float (&v)[16], bool live, uint4& lo, uint4& hi) {
  lo = make_uint4(pack_h2(v[0], v[1]), pack_h2(v[2], v[3]), pack_h2(v[4], v[5]), pack_h2(v[6], v[7]));
  hi = make_uint4(pack_h2(v[8], v[9]), pack_h2(v[10], v[11]), pack_h2(v[12], v[13]), pack_h2(v[14], v[15]));
  if (!live) { lo = make_uint4(0, 0, 0, 0); hi = lo; }
}
__device__ __forceinline__ uint32_t hmax2u(uint32_t a, uint32_t b) {
  __half2 r = __hmax2(*reinterpret_cast<__half2*>(&a), *reinterpret_cast<__half2*>(&b));
  return *reinterpret_cast<uint32_t*>(&r);
}
// Rows of a window are the lanes of a warp and lanes 30 / 31 are dead rows that hold exact zeros, so a ROTATING
// shuffle implements both the zero padding of the tap shifts and the -inf padding of the max-pool (values >= 0).
__device__ __forceinline__ uint32_t pool3(uint32_t v, int lane) {
  const uint32_t up = __shfl_sync(0xffffffffu, v, (lane + 31) & 31), dn = __shfl_sync(0xffffffffu, v, (lane + 1) & 31);
  return hmax2u(v, hmax2u(up, dn));
}
__device__ __forceinline__ uint4 pool3x4(uint4 v, int lane) {
  return make_uint4(pool3(v.x, lane), pool3(v.y, lane), pool3(v.z, lane), pool3(v.w, lane));
}
__device__ __forceinline__ float shf(float v, int src_lane) { return __shfl_sync(0xffffffffu, v, src_lane); }
// ReLU + fp16 pack of 16 accumulator columns without bias / dropout: convert first, clamp on packed halves
__device__ __forceinline__ void relu_pack16(const float (&v)[16], bool live, uint4& lo, uint4& hi) {
  uint32_t r[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) r[j] = live ? hmax2u(pack_h2(v[2 * j], v[2 * j + 1]), 0u) : 0u;
  lo = make_uint4(r[0], r[1], r[2], r[3]);
  hi = make_uint4(r[4], r[5], r[6], r[7]);
}

template <bool DROP>
__global__ void __launch_bounds__(512, 1) tc_conv_kernel(const ConvArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  const uint32_t sbase = smem_u32(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int grp = tid >> 8, gt = tid & 255;
  unsigned char* reg = smem + grp * G_BYTES;          // this group's activation region
  __half* raw = reinterpret_cast<__half*>(smem + OFF_RAW + grp * RAW_BYTES);  // raw fp16 copy of the tile's 4 windows
  const uint32_t rbase = sbase + grp * G_BYTES;
  const uint32_t gbar = sbase + OFF_BAR + grp * 24;   // phase barriers A, B, C of the group
  const uint32_t wbar = sbase + OFF_BAR + 48;         // weight barrier (CTA-wide)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_BAR + 64);
  volatile int* abort_flag = reinterpret_cast<volatile int*>(smem + OFF_BAR + 72);
  const float* sbias = reinterpret_cast<const float*>(smem + OFF_W + WI_BIAS);

  for (int i = tid; i < OFF_W / 16; i += 512) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    *abort_flag = 0;
    for (int i = 0; i < 7; ++i) mbar_init(sbase + OFF_BAR + 8 * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(smem_u32(tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot + grp * 256;

  const int npair = (a.ntile4 + 1) >> 1;
  const long long total = (long long)a.S * npair;
  const long long per = (total + gridDim.x - 1) / gridDim.x;
  const long long beg = per * blockIdx.x, end = min(total, beg + per);
  const int row = gt & 127, half = gt >> 7;
  const int wq = row >> 5, t = row & 31;  // t == lane
  const uint32_t lane_addr = tmem + ((uint32_t)(row & ~31) << 16);
  const uint32_t rowoff = (uint32_t)(ROW0 + row) * 16;

  const uint64_t dX = umma_desc(rbase + R_X + ROW0 * 16, CS, 128);
  const uint64_t dXP = umma_desc(rbase + R_XP + ROW0 * 16, CS, 128);
  const uint64_t dM1 = umma_desc(rbase + R_M1 + ROW0 * 16, CS, 128);
  const uint64_t dM1P = umma_desc(rbase + R_M1P + ROW0 * 16, CS, 128);
  const uint64_t dT2 = umma_desc(rbase + R_T2 + ROW0 * 16, CS, 128);
  const uint64_t dT3 = umma_desc(rbase + R_T3 + ROW0 * 16, CS, 128);
  const uint64_t dWB1 = umma_desc(sbase + OFF_W + WI_B1, 2304, 128);
  const uint64_t dWB4 = umma_desc(sbase + OFF_W + WI_B4, 512, 128);
  const uint64_t dWC2 = umma_desc(sbase + OFF_W + WI_B2B, 48 * 16, 128);
  const uint64_t dWC3 = umma_desc(sbase + OFF_W + WI_B3B, 80 * 16, 128);

  // wait for an MMA phase: one warp polls the mbarrier, the others block on the group's named barrier
  auto phase_wait = [&](uint32_t bar, uint32_t parity, int code) {
    if ((gt >> 5) == 0) mbar_wait(bar, parity, a.status, code, abort_flag);
    group_sync(grp);
    tc_fence_after();
  };
  // raw windows of a tile: global fp32 -> registers (9 floats / thread, coalesced) -> fp16 staging buffer
  float pre[9];
  auto prefetch = [&](long long item) {
    const bool valid = item < end;
    const int ptile = valid ? (int)(item % npair) * 2 + grp : 0;
#pragma unroll
    for (int j = 0; j < 9; ++j) {
      const int idx = gt + 256 * j;
      const int gw = ptile * 4 + idx / 540;
      const bool okx = valid && idx < 2160 && ptile < a.ntile4 && gw < a.B;
      pre[j] = okx ? __ldg(a.x + (long long)ptile * 2160 + idx) : 0.f;
    }
  };
  auto stash = [&]() {
#pragma unroll
    for (int j = 0; j < 9; ++j) {
      const int idx = gt + 256 * j;
      if (idx < 2160) raw[idx] = __float2half_rn(pre[j]);
    }
  };

  int cur_s = -1;
  uint32_t ph = 0, wph = 0;
  prefetch(beg);
  stash();
  group_sync(grp);

  for (long long it = beg; it < end; ++it) {
    const int s = (int)(it / npair);
    const int tile = (int)(it % npair) * 2 + grp;  // an odd tile count gives group 1 a dummy (all-dead) tile
    if (s != cur_s) {  // stage this sample's conv weights; every MMA that read the old ones has completed
      cur_s = s;
      __syncthreads();
      if (tid == 0) {
        const unsigned char* src = a.blob + (long long)s * a.blob_stride;
        mbar_expect_tx(wbar, CONV_IMG);
        for (int o = 0; o < CONV_IMG; o += 16384) bulk_g2s(sbase + OFF_W + o, src + o, min(16384, CONV_IMG - o), wbar);
      }
      mbar_wait(wbar, wph, a.status, 1, abort_flag);
      wph ^= 1;
    }
    // ---- X / XP: 8-feature fp16 chunks of the windows and of their MaxPool1d(3,1,1) (-inf padding) ----
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int u = gt + 256 * i, r = u & 127, c = u >> 7;  // i = 0: chunks 0,1; i = 1: chunks 2,3
      const int tt = r & 31, w4 = r >> 5, gw = tile * 4 + w4;
      const bool lv = tt < 30 && gw < a.B && tile < a.ntile4;
      uint32_t f[4] = {0u, 0u, 0u, 0u}, pm[4] = {0u, 0u, 0u, 0u};
      if (lv && c < 3) {
        const uint32_t* pr = reinterpret_cast<const uint32_t*>(raw + w4 * 540 + tt * 18 + c * 8);  // 4-byte aligned
        const int nw = c < 2 ? 4 : 1;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (j < nw) {
            const uint32_t v = pr[j];
            uint32_t m = v;
            if (tt > 0) m = hmax2u(m, pr[j - 9]);
            if (tt < 29) m = hmax2u(m, pr[j + 9]);
            f[j] = v;
            pm[j] = m;
          }
        }
        if (c == 2) f[1] = pm[1] = 0x00003C00u;  // feature 18 = 1.0 (fp16), feature 19 = 0: carries the biases
      }
      *reinterpret_cast<uint4*>(reg + R_X + c * CS + (ROW0 + r) * 16) = make_uint4(f[0], f[1], f[2], f[3]);
      *reinterpret_cast<uint4*>(reg + R_XP + c * CS + (ROW0 + r) * 16) = make_uint4(pm[0], pm[1], pm[2], pm[3]);
    }
    fence_async_smem();
    tc_fence_before();
    group_sync(grp);
    // ---- phase A: module 1.  TMEM cols: conv5 0..31 | conv3 32..63 | conv1 64..95 | convpool 96..127 ----
    if (gt == 0) {
      tc_fence_after();
      constexpr int shs[5] = {0, -1, 1, -2, 2}, nsh[5] = {96, 64, 64, 32, 32}, osh[5] = {0, 6144, 10240, 14336, 16384};
#pragma unroll
      for (int q = 0; q < 5; ++q)
#pragma unroll
        for (int ks = 0; ks < 2; ++ks)
          umma(tmem, dX + (uint64_t)((2 * ks * CS + shs[q] * 16) >> 4),
               umma_desc(sbase + OFF_W + WI_A + osh[q] + 2 * ks * nsh[q] * 16, nsh[q] * 16, 128), umma_idesc(nsh[q]), (q | ks) != 0);
#pragma unroll
      for (int tp = 0; tp < 3; ++tp)
#pragma unroll
        for (int ks = 0; ks < 2; ++ks)
          umma(tmem + 96, dXP + (uint64_t)((2 * ks * CS + (tp - 1) * 16) >> 4),
               umma_desc(sbase + OFF_W + WI_A + 18432 + tp * 2048 + 2 * ks * 512, 512, 128), umma_idesc(32), (tp | ks) != 0);
      umma_commit(gbar);
    }
    prefetch(it + 1);  // next tile's windows: global loads stay in flight for the rest of this tile
    phase_wait(gbar, ph, 2);
    const int gw = tile * 4 + wq;
    const bool live = t < 30 && gw < a.B && tile < a.ntile4;
#pragma unroll
    for (int bp = 0; bp < 2; ++bp) {
      float acc[2][16];
      tmem_ld16(lane_addr + (2 * bp) * 32 + half * 16, acc[0]);
      tmem_ld16(lane_addr + (2 * bp + 1) * 32 + half * 16, acc[1]);
      tmem_ld_wait();
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int cb = 2 * bp + q;                         // TMEM column block
        const int br = cb == 0 ? 2 : cb == 1 ? 1 : cb == 2 ? 0 : 3;  // -> M1 channel group / dropout layer
        uint4 lo, hi;
        if (DROP) {
          act16<DROP, false>(acc[q], nullptr, live, a, br, s, gw, t, half * 16, 27);
          pack16(acc[q], live, lo, hi);
        } else {
          relu_pack16(acc[q], live, lo, hi);
        }
        unsigned char* dst = reg + R_M1 + (br * 4 + half * 2) * CS + rowoff;
        *reinterpret_cast<uint4*>(dst) = lo;
        *reinterpret_cast<uint4*>(dst + CS) = hi;
        *reinterpret_cast<uint4*>(dst + R_M1P) = pool3x4(lo, lane);  // R_M1P - R_M1 == 16 chunks
        *reinterpret_cast<uint4*>(dst + R_M1P + CS) = pool3x4(hi, lane);
      }
    }
    fence_async_smem();
    tc_fence_before();
    group_sync(grp);
    // ---- phase B: module-2 1x1 convs: [b1 | b2a | b3a] (N = 144) on M1, b4 (N = 32) on pooled M1 ----
    if (gt == 0) {
      tc_fence_after();
#pragma unroll
      for (int ks = 0; ks < 8; ++ks)
        umma(tmem, dM1 + (uint64_t)((2 * ks * CS) >> 4), dWB1 + (uint64_t)((2 * ks * 2304) >> 4), umma_idesc(144), ks != 0);
#pragma unroll
      for (int ks = 0; ks < 8; ++ks)
        umma(tmem + 144, dM1P + (uint64_t)((2 * ks * CS) >> 4), dWB4 + (uint64_t)((2 * ks * 512) >> 4), umma_idesc(32), ks != 0);
      umma_commit(gbar + 8);
    }
    phase_wait(gbar + 8, ph, 3);
    unsigned char* frow = a.feat + ((long long)s * a.ntile128 + (gw >> 7)) * FEAT_TILE_BYTES + (long long)(gw & 127) * 16 +
                          (long long)t * 10 * 2048;
    // 11 groups of 16 columns: g0 = b1 -> feat ch 0..15 | g1..4 = b2a -> T2 | g5..8 = b3a -> T3 | g9,10 = b4 -> feat ch 48..79
#pragma unroll
    for (int gi = 0; gi < 3; ++gi) {
      const int g0 = half + 4 * gi, g1 = g0 + 2;  // half 0: (0,2)(4,6)(8,10); half 1: (1,3)(5,7)(9,-)
      const bool two = g1 < 11;
      float acc[2][16];
      tmem_ld16(lane_addr + g0 * 16, acc[0]);
      if (two) tmem_ld16(lane_addr + g1 * 16, acc[1]);
      tmem_ld_wait();
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int g = q ? g1 : g0;
        if (q && !two) break;
        uint4 lo, hi;
        if (g == 0) {
          if (DROP) { act16<DROP, false>(acc[q], nullptr, live, a, 4, s, gw, t, 0, 16); pack16(acc[q], live, lo, hi); }
          else relu_pack16(acc[q], live, lo, hi);
          if (live) { *reinterpret_cast<uint4*>(frow) = lo; *reinterpret_cast<uint4*>(frow + 2048) = hi; }
        } else if (g < 9) {
          relu_pack16(acc[q], live, lo, hi);
          unsigned char* dst = reg + (g < 5 ? R_T2 + (g - 1) * 2 * CS : R_T3 + (g - 5) * 2 * CS) + rowoff;
          *reinterpret_cast<uint4*>(dst) = lo;
          *reinterpret_cast<uint4*>(dst + CS) = hi;
        } else {
          if (DROP) { act16<DROP, false>(acc[q], nullptr, live, a, 9, s, gw, t, (g - 9) * 16, 32); pack16(acc[q], live, lo, hi); }
          else relu_pack16(acc[q], live, lo, hi);
          if (live) {
            *reinterpret_cast<uint4*>(frow + (6 + (g - 9) * 2) * 2048) = lo;
            *reinterpret_cast<uint4*>(frow + (7 + (g - 9) * 2) * 2048) = hi;
          }
        }
      }
    }
    fence_async_smem();
    tc_fence_before();
    group_sync(grp);
    // ---- phase C: b2b (k3 over T2) / b3b (k5 over T3): taps concatenated along N, shifts applied in the epilogue ----
    if (gt == 0) {
      tc_fence_after();
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
        umma(tmem, dT2 + (uint64_t)((2 * ks * CS) >> 4), dWC2 + (uint64_t)((2 * ks * 48 * 16) >> 4), umma_idesc(48), ks != 0);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
        umma(tmem + 48, dT3 + (uint64_t)((2 * ks * CS) >> 4), dWC3 + (uint64_t)((2 * ks * 80 * 16) >> 4), umma_idesc(80), ks != 0);
      umma_commit(gbar + 16);
    }
    stash();  // the prefetched windows have landed long ago; the barrier inside phase_wait publishes them
    phase_wait(gbar + 16, ph, 4);
    {
      const int lm1 = (lane + 31) & 31, lm2 = (lane + 30) & 31, lp1 = (lane + 1) & 31, lp2 = (lane + 2) & 31;
      float out[16];
      if (half == 0) {  // b2b: out[t] = P0[t-1] + P1[t] + P2[t+1]
        float p[3][16];
#pragma unroll
        for (int tp = 0; tp < 3; ++tp) tmem_ld16(lane_addr + tp * 16, p[tp]);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) out[j] = shf(p[0][j], lm1) + p[1][j] + shf(p[2][j], lp1);
      } else {  // b3b: out[t] = sum_tap P_tap[t + tap - 2]
        float p[3][16];
#pragma unroll
        for (int tp = 0; tp < 3; ++tp) tmem_ld16(lane_addr + 48 + tp * 16, p[tp]);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) out[j] = shf(p[0][j], lm2) + shf(p[1][j], lm1) + p[2][j];
        tmem_ld16(lane_addr + 48 + 48, p[0]);
        tmem_ld16(lane_addr + 48 + 64, p[1]);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) out[j] += shf(p[0][j], lp1) + shf(p[1][j], lp2);
      }
      act16<DROP, true>(out, sbias + 304 + half * 16, live, a, half ? 8 : 6, s, gw, t, 0, 16);
      uint4 lo, hi;
      pack16(out, live, lo, hi);
      if (live) {
        *reinterpret_cast<uint4*>(frow + (2 + half * 2) * 2048) = lo;
        *reinterpret_cast<uint4*>(frow + (3 + half * 2) * 2048) = hi;
      }
    }
    tc_fence_before();
    ph ^= 1;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(*tmem_slot, 512);
}
