// Probe: which shared-memory descriptor / instruction-descriptor settings make tcgen05.mma read the activation and weight
// IMAGES of the fused kernels (brl_tc_conv.cuh / brl_tc_train.cuh) as a TRANSPOSED operand, and in which formats.
//   image layout of an activation tile:   element (channel c, row r) at (c / 8) * CS + r * 16 + (c % 8) * 2,  CS = 132 * 16
//   image layout of a weight tile [N, C]: element (n, c)             at (c / 8) * N * 16 + n * 16 + (c % 8) * 2
// Cases (all values are small integers, so every product / sum is exact in fp16, bf16 and fp32):
//   1 forward-like    D[r, n] = sum_c  act[c, r + sh] * w[n, c]     A K-major, B K-major                     (known good)
//   2 weight gradient D[c, n] = sum_r  act[c, r + sh] * g[n, r]     A MN-major, B MN-major over 128 rows (8 MMAs)
//   3 input gradient  D[r, c] = sum_n  g[n, r + sh]   * w[n, c]     A K-major, B MN-major (the forward weight image, transposed)
// each with fp16 and bf16 operands.  (Measured on B200: kind 4 = case 1 with fp16 A and bf16 B in ONE instruction raises
// 'illegal instruction' -- kind::f16 wants both operands in the same format; pass {4, 0, 1, 0, 64} to see it.)
// Build (shared cudart, so no shipped artefact embeds driver API names) and run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 --cudart shared -o gpurun_out/umma_probe tools/ubench/umma_probe.cu && gpurun_out/umma_probe
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../bayesrul_b200/csrc/brl_tc_ptx.cuh"

using namespace brl;

constexpr int ROWS = 132, CS = ROWS * 16, ROW0 = 2;
constexpr int ACT_CH = 128, G_CH = 64;             // channels of the activation / gradient images
constexpr int OFF_ACT = 0, OFF_G = OFF_ACT + (ACT_CH / 8) * CS, OFF_W = OFF_G + (G_CH / 8) * CS;  // W: [N = 64][C = 128]
constexpr int W_N = 64;
constexpr int OFF_END = OFF_W + (ACT_CH / 8) * W_N * 16;
constexpr int SMEM = OFF_END + 64;

// idesc: fp32 D, formats fa / fb (0 fp16, 1 bf16), majors (1 = MN-major), M = 128
__host__ __device__ constexpr uint32_t idesc(int n, int fa, int fb, int amn, int bmn) {
  return (1u << 4) | ((uint32_t)fa << 7) | ((uint32_t)fb << 10) | ((uint32_t)amn << 15) | ((uint32_t)bmn << 16) | ((uint32_t)(n >> 3) << 17) |
         ((128u >> 4) << 24);
}

struct Case { int kind, fa, fb, shift, n; };

__global__ void __launch_bounds__(128, 1) probe(const unsigned char* img, float* out, Case cs) {
  extern __shared__ __align__(128) unsigned char smem[];
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar = sbase + OFF_END;
  uint32_t* slot = reinterpret_cast<uint32_t*>(smem + OFF_END + 16);
  for (int i = threadIdx.x; i < OFF_END / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = reinterpret_cast<const uint4*>(img)[i];
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) tmem_alloc(smem_u32(slot), 128);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x < 32) {
    if (elect_one()) {
      if (cs.kind == 1 || cs.kind == 4) {  // K = 16 channels x 8 k-steps = 128 channels
        for (int ks = 0; ks < 8; ++ks)
          umma(tmem, umma_desc(sbase + OFF_ACT + 2 * ks * CS + (ROW0 + cs.shift) * 16, CS, 128),
               umma_desc(sbase + OFF_W + 2 * ks * W_N * 16, W_N * 16, 128), idesc(cs.n, cs.fa, cs.fb, 0, 0), ks != 0);
      } else if (cs.kind == 2) {  // K = rows: 16 per MMA, 8 MMAs; A: M = 128 channels (SBO = CS), B: N = n channels of g
        for (int ks = 0; ks < 8; ++ks)
          umma(tmem, umma_desc(sbase + OFF_ACT + (ROW0 + cs.shift + 16 * ks) * 16, 128, CS),
               umma_desc(sbase + OFF_G + (ROW0 + 16 * ks) * 16, 128, CS), idesc(cs.n, cs.fa, cs.fb, 1, 1), ks != 0);
      } else {  // kind 3: K = n (64 gradient channels: 4 k-steps); A = g K-major; B = w image read as [N' = c][K' = n]
        for (int ks = 0; ks < 4; ++ks)
          umma(tmem, umma_desc(sbase + OFF_G + 2 * ks * CS + (ROW0 + cs.shift) * 16, CS, 128),
               umma_desc(sbase + OFF_W + ks * 256, 128, W_N * 16), idesc(cs.n, cs.fa, cs.fb, 0, 1), ks != 0);
      }
      umma_commit(bar);
    }
    __syncwarp();
  }
  mbar_wait(bar, 0, reinterpret_cast<int*>(out + 128 * 128), 1);
  tc_fence_after();
  const int row = threadIdx.x;
  const uint32_t la = tmem + ((uint32_t)(row & ~31) << 16);
  for (int c0 = 0; c0 < cs.n; c0 += 16) {
    float v[16];
    tmem_ld16(la + c0, v);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) out[row * 128 + c0 + j] = v[j];
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 128);
}

static uint16_t enc(int v, int fmt) {
  if (fmt == 0) { __half h = __float2half((float)v); return *reinterpret_cast<uint16_t*>(&h); }
  __nv_bfloat16 b = __float2bfloat16((float)v);
  return *reinterpret_cast<uint16_t*>(&b);
}

int main() {
  std::vector<int> act(ACT_CH * ROWS), g(G_CH * ROWS), w(W_N * ACT_CH);
  srand(7);
  for (auto& v : act) v = rand() % 7 - 3;
  for (auto& v : g) v = rand() % 5 - 2;
  for (auto& v : w) v = rand() % 5 - 2;
  unsigned char* dimg;
  float* dout;
  cudaMalloc(&dimg, OFF_END);
  cudaMalloc(&dout, 128 * 128 * 4 + 16);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
  const Case cases[] = {{1, 0, 0, 0, 64},  {1, 0, 0, -1, 64}, {1, 1, 1, 1, 64},
                        {2, 0, 0, 0, 64},  {2, 0, 0, 1, 64},  {2, 1, 1, -2, 32}, {2, 1, 1, 0, 16}, {3, 0, 0, 0, 128},
                        {3, 1, 1, -1, 128}, {3, 1, 1, 2, 112}};
  int bad = 0;
  for (const Case& cs : cases) {
    std::vector<uint16_t> img(OFF_END / 2, 0);
    auto put = [&](int off, int v, int fmt) { img[off / 2] = enc(v, fmt); };
    // activation image in format fa (kind 3 does not use it), gradient image: fb for kind 2, fa for kind 3; weights: fb
    for (int c = 0; c < ACT_CH; ++c)
      for (int r = 0; r < ROWS; ++r) put(OFF_ACT + (c / 8) * CS + r * 16 + (c % 8) * 2, act[c * ROWS + r], cs.fa);
    for (int c = 0; c < G_CH; ++c)
      for (int r = 0; r < ROWS; ++r) put(OFF_G + (c / 8) * CS + r * 16 + (c % 8) * 2, g[c * ROWS + r], cs.kind == 2 ? cs.fb : cs.fa);
    for (int n = 0; n < W_N; ++n)
      for (int c = 0; c < ACT_CH; ++c) put(OFF_W + (c / 8) * W_N * 16 + n * 16 + (c % 8) * 2, w[n * ACT_CH + c], cs.fb);
    cudaMemcpy(dimg, img.data(), OFF_END, cudaMemcpyHostToDevice);
    cudaMemset(dout, 0, 128 * 128 * 4 + 16);
    probe<<<1, 128, SMEM>>>(dimg, dout, cs);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<float> out(128 * 128);
    cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0;
    for (int i = 0; i < 128; ++i)
      for (int j = 0; j < cs.n; ++j) {
        double ref = 0;
        if (cs.kind == 1 || cs.kind == 4) for (int c = 0; c < ACT_CH; ++c) ref += act[c * ROWS + ROW0 + i + cs.shift] * w[j * ACT_CH + c];
        else if (cs.kind == 2) for (int r = 0; r < 128; ++r) ref += act[i * ROWS + ROW0 + r + cs.shift] * g[j * ROWS + ROW0 + r];
        else for (int n = 0; n < G_CH; ++n) ref += g[n * ROWS + ROW0 + i + cs.shift] * w[n * ACT_CH + j];
        maxerr = std::max(maxerr, std::abs(ref - out[i * 128 + j]));
      }
    printf("case kind=%d fa=%d fb=%d shift=%+d n=%3d : %s  max|err| = %g  (cuda: %s)\n", cs.kind, cs.fa, cs.fb, cs.shift, cs.n,
           maxerr == 0 ? "OK  " : "FAIL", maxerr, cudaGetErrorString(e));
    bad += maxerr != 0;
    if (e != cudaSuccess) return 2;
  }
  printf("%d failing case(s)\n", bad);
  return 0;
}
