// tcgen05 conv-stack kernel (included by brl_tc.cu after the PTX wrappers and geometry constants).
//
// One persistent CTA per SM, 512 threads = two independent 256-thread groups.  Each group owns one 128-row tile
// (4 windows x 32 rows) at a time, its own 66 KB activation region, 256 TMEM columns, three mbarriers and one
// named barrier, so while one group sits in an MMA wait / barrier the other runs its epilogue: the tensor pipe and
// the issue slots see two tiles in flight.  The fp16 weights of the current MC sample (86 KB) are shared.
//
// Activation region of a group (K-major, no swizzle: [8-channel chunk][132 rows][16 B]):
//   M1  16 chunks  module-1 output, 4 branches x 32 channels (27 real + const-1 channel + zeros)
//   M1P 16 chunks  MaxPool1d(3,1,1)(M1)
//   T2 / T3 (8 + 8 chunks) alias M1   (written after the MMAs reading M1 have completed)
//   X / XP  (4 + 4 chunks) alias M1P  (read by phase A only, M1P is written after phase A completed)
// Biases ride in the MMAs: input feature 18 is a constant 1 whose weight row (centre tap only) holds the bias;
// conv1's column 27 reproduces the constant into M1 (and, pooled, into M1P) for the module-2 1x1 convs.
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

// 16 accumulator columns -> (+bias), ReLU, (dropout), fp16; zeros for dead rows
template <bool DROP, bool BIAS>
__device__ __forceinline__ void epi16(const float (&acc)[16], const float* bias, bool live, uint4& lo, uint4& hi,
                                      const ConvArgs& a, int layer, int s, int gw, int t, int ch0, int nvalid) {
  float v[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    float u = fmaxf(BIAS ? acc[j] + bias[j] : acc[j], 0.f);
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
  lo = make_uint4(pack_h2(v[0], v[1]), pack_h2(v[2], v[3]), pack_h2(v[4], v[5]), pack_h2(v[6], v[7]));
  hi = make_uint4(pack_h2(v[8], v[9]), pack_h2(v[10], v[11]), pack_h2(v[12], v[13]), pack_h2(v[14], v[15]));
  if (!live) { lo = make_uint4(0, 0, 0, 0); hi = lo; }
}

__device__ __forceinline__ void fmax8(float (&d)[8], const float (&s)[8]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) d[j] = fmaxf(d[j], s[j]);
}

template <bool DROP>
__global__ void __launch_bounds__(512, 1) tc_conv_kernel(const ConvArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  const uint32_t sbase = smem_u32(smem);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int grp = tid >> 8, gt = tid & 255;
  unsigned char* reg = smem + grp * G_BYTES;          // this group's activation region
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
  const int wq = row >> 5, t = row & 31;
  const uint32_t lane_addr = tmem + ((uint32_t)(row & ~31) << 16);
  const uint32_t rowoff = (uint32_t)(ROW0 + row) * 16;

  // descriptor bases (all operand addresses are compile-time offsets from these)
  const uint64_t dX = umma_desc(rbase + R_X + ROW0 * 16, CS, 128);
  const uint64_t dXP = umma_desc(rbase + R_XP + ROW0 * 16, CS, 128);
  const uint64_t dM1 = umma_desc(rbase + R_M1 + ROW0 * 16, CS, 128);
  const uint64_t dM1P = umma_desc(rbase + R_M1P + ROW0 * 16, CS, 128);
  const uint64_t dT2 = umma_desc(rbase + R_T2 + ROW0 * 16, CS, 128);
  const uint64_t dT3 = umma_desc(rbase + R_T3 + ROW0 * 16, CS, 128);
  const uint64_t dWA = umma_desc(sbase + OFF_W + WI_A, 512, 128);
  const uint64_t dWB1 = umma_desc(sbase + OFF_W + WI_B1, 2304, 128);
  const uint64_t dWB4 = umma_desc(sbase + OFF_W + WI_B4, 512, 128);
  const uint64_t dWC2 = umma_desc(sbase + OFF_W + WI_B2B, 256, 128);
  const uint64_t dWC3 = umma_desc(sbase + OFF_W + WI_B3B, 256, 128);

  int cur_s = -1;
  uint32_t ph = 0, wph = 0;
  bool ok = true;
  (void)ok;

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
      ok = mbar_wait(wbar, wph, a.status, 1, abort_flag);
      wph ^= 1;
    }
    // ---- stage the 4 windows (fp32 -> fp16, 8-feature chunks) and their MaxPool1d(3,1,1) (-inf padding) ----
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int u = gt + 256 * i, r = u & 127, c = u >> 7;  // i = 0: chunks 0,1; i = 1: chunks 2,3
      const int tt = r & 31, gw = tile * 4 + (r >> 5);
      const bool lv = tt < 30 && gw < a.B && tile < a.ntile4;
      float f[8], pm[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = pm[j] = 0.f;
      if (lv && c < 3) {
        const float* px = a.x + (long long)gw * 540 + tt * 18 + c * 8;
        const int nf = c < 2 ? 8 : 2;
        float lo[8], hi[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) lo[j] = hi[j] = -INFINITY;
        for (int j = 0; j < nf; j += 2) {
          const float2 p2 = __ldg(reinterpret_cast<const float2*>(px + j));
          f[j] = p2.x; f[j + 1] = p2.y;
          if (tt > 0) { const float2 q = __ldg(reinterpret_cast<const float2*>(px - 18 + j)); lo[j] = q.x; lo[j + 1] = q.y; }
          if (tt < 29) { const float2 q = __ldg(reinterpret_cast<const float2*>(px + 18 + j)); hi[j] = q.x; hi[j + 1] = q.y; }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) pm[j] = j < nf ? fmaxf(f[j], fmaxf(lo[j], hi[j])) : 0.f;
        if (c == 2) f[2] = pm[2] = 1.0f;  // constant-1 feature carrying the biases
      }
      *reinterpret_cast<uint4*>(reg + R_X + c * CS + (ROW0 + r) * 16) =
          make_uint4(pack_h2(f[0], f[1]), pack_h2(f[2], f[3]), pack_h2(f[4], f[5]), pack_h2(f[6], f[7]));
      *reinterpret_cast<uint4*>(reg + R_XP + c * CS + (ROW0 + r) * 16) =
          make_uint4(pack_h2(pm[0], pm[1]), pack_h2(pm[2], pm[3]), pack_h2(pm[4], pm[5]), pack_h2(pm[6], pm[7]));
    }
    fence_async_smem();
    tc_fence_before();
    group_sync(grp);
    // ---- phase A: module 1, 4 branches x (N = 32), taps = row-shifted A descriptors ----
    if (gt == 0) {
      tc_fence_after();
      constexpr uint32_t idA = umma_idesc(32);
      constexpr int ntap[4] = {1, 3, 5, 3}, tap0[4] = {0, 1, 4, 9}, pad[4] = {0, 1, 2, 1};
#pragma unroll
      for (int br = 0; br < 4; ++br)
#pragma unroll
        for (int tp = 0; tp < ntap[br]; ++tp)
#pragma unroll
          for (int ks = 0; ks < 2; ++ks)
            umma(tmem + br * 32, (br == 3 ? dXP : dX) + (uint64_t)((2 * ks * CS + (tp - pad[br]) * 16) >> 4),
                 dWA + (uint64_t)(((tap0[br] + tp) * WA_TAP + 2 * ks * 512) >> 4), idA, (tp | ks) != 0);
      umma_commit(gbar);
    }
    ok = mbar_wait(gbar, ph, a.status, 2, abort_flag);
    tc_fence_after();
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
        const int br = 2 * bp + q;
        uint4 lo, hi;
        epi16<DROP, false>(acc[q], nullptr, live, lo, hi, a, br, s, gw, t, half * 16, 27);
        unsigned char* dst = reg + R_M1 + (br * 4 + half * 2) * CS + rowoff;
        *reinterpret_cast<uint4*>(dst) = lo;
        *reinterpret_cast<uint4*>(dst + CS) = hi;
      }
    }
    tc_fence_before();
    group_sync(grp);
    // ---- MaxPool1d(3,1,1) of module-1 output (post-ReLU >= 0, zero pad rows act as -inf) ----
#pragma unroll 2
    for (int u = gt; u < 2048; u += 256) {
      const int r = u & 127, c = u >> 7;
      const unsigned char* p = reg + R_M1 + c * CS + (ROW0 + r) * 16;
      const uint4 v = hmax4(hmax4(*reinterpret_cast<const uint4*>(p - 16), *reinterpret_cast<const uint4*>(p)),
                            *reinterpret_cast<const uint4*>(p + 16));
      *reinterpret_cast<uint4*>(reg + R_M1P + c * CS + (ROW0 + r) * 16) = v;
    }
    fence_async_smem();
    group_sync(grp);
    // ---- phase B: module-2 1x1 convs: [b1 | b2a | b3a] (N = 144) on M1, b4 (N = 32) on pooled M1 ----
    if (gt == 0) {
      tc_fence_after();
      constexpr uint32_t idB1 = umma_idesc(144), idB4 = umma_idesc(32);
#pragma unroll
      for (int ks = 0; ks < 8; ++ks)
        umma(tmem, dM1 + (uint64_t)((2 * ks * CS) >> 4), dWB1 + (uint64_t)((2 * ks * 2304) >> 4), idB1, ks != 0);
#pragma unroll
      for (int ks = 0; ks < 8; ++ks)
        umma(tmem + 144, dM1P + (uint64_t)((2 * ks * CS) >> 4), dWB4 + (uint64_t)((2 * ks * 512) >> 4), idB4, ks != 0);
      umma_commit(gbar + 8);
    }
    ok = mbar_wait(gbar + 8, ph, a.status, 3, abort_flag);
    tc_fence_after();
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
          epi16<DROP, false>(acc[q], nullptr, live, lo, hi, a, 4, s, gw, t, 0, 16);
          if (live) { *reinterpret_cast<uint4*>(frow) = lo; *reinterpret_cast<uint4*>(frow + 2048) = hi; }
        } else if (g < 9) {
          epi16<false, false>(acc[q], nullptr, live, lo, hi, a, 0, s, gw, t, 0, 16);
          unsigned char* dst = reg + (g < 5 ? R_T2 + (g - 1) * 2 * CS : R_T3 + (g - 5) * 2 * CS) + rowoff;
          *reinterpret_cast<uint4*>(dst) = lo;
          *reinterpret_cast<uint4*>(dst + CS) = hi;
        } else {
          epi16<DROP, false>(acc[q], nullptr, live, lo, hi, a, 9, s, gw, t, (g - 9) * 16, 32);
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
    // ---- phase C: b2b (k3 over T2) and b3b (k5 over T3), N = 16 each ----
    if (gt == 0) {
      tc_fence_after();
      constexpr uint32_t idC = umma_idesc(16);
#pragma unroll
      for (int tp = 0; tp < 3; ++tp)
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          umma(tmem + 176, dT2 + (uint64_t)((2 * ks * CS + (tp - 1) * 16) >> 4), dWC2 + (uint64_t)((tp * WC_TAP + 2 * ks * 256) >> 4),
               idC, (tp | ks) != 0);
#pragma unroll
      for (int tp = 0; tp < 5; ++tp)
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          umma(tmem + 192, dT3 + (uint64_t)((2 * ks * CS + (tp - 2) * 16) >> 4), dWC3 + (uint64_t)((tp * WC_TAP + 2 * ks * 256) >> 4),
               idC, (tp | ks) != 0);
      umma_commit(gbar + 16);
    }
    ok = mbar_wait(gbar + 16, ph, a.status, 4, abort_flag);
    tc_fence_after();
    {
      float acc[16];
      tmem_ld16(lane_addr + 176 + half * 16, acc);
      tmem_ld_wait();
      uint4 lo, hi;
      epi16<DROP, true>(acc, sbias + 304 + half * 16, live, lo, hi, a, half ? 8 : 6, s, gw, t, 0, 16);
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
