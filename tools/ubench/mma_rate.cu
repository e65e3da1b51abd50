// Microbenchmark: issue/execute cost of small tcgen05.mma (M=128, K=16, kind::f16) on one SM as a function of N,
// operand source (A in shared memory, no-swizzle K-major, vs A in TMEM) and accumulator dependency.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu && ./mma_rate
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
__host__ __device__ constexpr uint32_t umma_idesc(int n) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24); }
__device__ __forceinline__ void umma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

// mode 0: SS same accumulator; 1: SS, NACC independent accumulators round-robin; 2: TS same accumulator; 3: TS round-robin
template <int N, int MODE, int NACC>
__global__ void __launch_bounds__(128, 1) k(long long* out, int reps) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const uint32_t sbase = smem_u32(smem);
  for (int i = threadIdx.x; i < 65536 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  const uint32_t b = smem_u32(&bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot;
  if (threadIdx.x < 32) {
    long long t0 = 0, t1 = 0, t2 = 0;
    const uint64_t dA = umma_desc(sbase, 2112, 128), dB = umma_desc(sbase + 32768, N * 16, 128);
    if (elect_one()) {
      t0 = clock64();
#pragma unroll 1
      for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const uint32_t d = tmem + ((MODE & 1) ? (i % NACC) * (256 / NACC >= N ? 256 / NACC : N) : 0);
          if (MODE < 2) umma_ss(d, dA + (uint64_t)(i * 264), dB + (uint64_t)((i & 1) * 2 * N), umma_idesc(N), 1u);
          else umma_ts(d, tmem + 256 + 8 * (i & 7), dB + (uint64_t)((i & 1) * 2 * N), umma_idesc(N), 1u);
        }
      }
      t1 = clock64();
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(b) : "memory");
    }
    __syncwarp();
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(b), "r"(0u) : "memory");
    }
    t2 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    // elected lane may not be lane 0: reduce through shuffles
    long long a0 = t1 - t0, a1 = t2 - t0;
    for (int o = 16; o; o >>= 1) {
      a0 = max(a0, __shfl_xor_sync(0xffffffffu, a0, o));
    }
    if (threadIdx.x == 0 && blockIdx.x == 0) { out[0] = a0; out[1] = a1; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

template <int N, int MODE, int NACC>
void run(long long* d, const char* name) {
  const int reps = 64;
  cudaFuncSetAttribute(k<N, MODE, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 8192);
  for (int w = 0; w < 2; ++w) k<N, MODE, NACC><<<1, 128, 65536 + 8192>>>(d, reps);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("%-28s N=%3d  issue %7.1f cyc/MMA   complete %7.1f cyc/MMA   (%s)\n", name, N, h[0] / (16.0 * reps), h[1] / (16.0 * reps),
         e == cudaSuccess ? "ok" : cudaGetErrorString(e));
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  run<16, 0, 1>(d, "SS same-acc");
  run<32, 0, 1>(d, "SS same-acc");
  run<48, 0, 1>(d, "SS same-acc");
  run<64, 0, 1>(d, "SS same-acc");
  run<96, 0, 1>(d, "SS same-acc");
  run<128, 0, 1>(d, "SS same-acc");
  run<144, 0, 1>(d, "SS same-acc");
  run<256, 0, 1>(d, "SS same-acc");
  run<32, 1, 4>(d, "SS 4 accumulators");
  run<64, 1, 4>(d, "SS 4 accumulators");
  run<128, 1, 2>(d, "SS 2 accumulators");
  run<16, 2, 1>(d, "TS (A in TMEM) same-acc");
  run<32, 2, 1>(d, "TS (A in TMEM) same-acc");
  run<64, 2, 1>(d, "TS (A in TMEM) same-acc");
  run<128, 2, 1>(d, "TS (A in TMEM) same-acc");
  run<144, 2, 1>(d, "TS (A in TMEM) same-acc");
  run<256, 2, 1>(d, "TS (A in TMEM) same-acc");
  run<32, 3, 4>(d, "TS 4 accumulators");
  run<64, 3, 4>(d, "TS 4 accumulators");
  return 0;
}
