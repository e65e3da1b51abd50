// Level-fused tcgen05 TRAINING kernels of the Inception conv stack (LRT / Flipout ELBO step) -- interface used by brl_api.cu.
// Reference semantics: tyxe.poutine.local_reparameterization / flipout around nets/inception.py:54-61,125-132 under
// svi.step (bayesian.py:146-147); see brl_tc_train.cu for the design.
#pragma once
#include <cstddef>
#include <cuda_runtime.h>

#include "brl_kernels.cuh"

namespace brl {

constexpr int TT_LAYERS = 10;  // conv layers 0..9 of the Inception table (brl_nets.cpp); layers 10 / 11 (fc, head) stay per-layer

// device buffers of one particle lane (carved from the caller's workspace)
struct TtLane {
  unsigned char* ximg;   // [ntile][X0 X1 X2 XP0 XP1 XP2][132 rows][16 B] fp16
  unsigned char* m1;     // [ntile][16 chunks][132][16 B] fp16: module-1 output, 4 branches x 32 channels (27 real)
  unsigned char* t2;     // [ntile][8 chunks] fp16
  unsigned char* t3;
  unsigned char* g[6];   // bf16 gradient images: [0..2] d/dM1 from b1 / b2a / b3a, [3] d/dMaxPool(M1) from b4, [4] d/dT2, [5] d/dT3
  float* rbuf;           // LRT: eps / (2 sd) of every conv layer output, [layer][ntile][NP][128 rows]
  unsigned char* blob;   // weight images of this step / particle
};
size_t tt_lane_bytes(long long B);
void tt_carve(unsigned char* base, long long B, TtLane& ln);

struct TtStep {
  const float* x;        // [B,30,18]
  long long B;
  int mode;              // BRL_MODE_LRT or BRL_MODE_FLIPOUT
  const float* mu;       // [P]
  const float* sigma;    // [P]  (LRT)
  const float* wsamp;    // [P]  (Flipout: the particle's weight draw)
  NoiseRef eps[TT_LAYERS];          // LRT eps streams (injected tensor [B, N*30] or Philox)
  const float* sgn_in[TT_LAYERS];   // Flipout [B, Cin]
  const float* sgn_out[TT_LAYERS];  // Flipout [B, Cout]
  float* feat;           // fp32 [B,80,30]: module-2 output = input buffer of the fc layer (per-layer engine)
  const float* feat_grad;  // fp32 [B,80,30]: its gradient (written by the fc layer's backward)
  float* g0;             // flat gradient accumulators [P] (brl_kernels.cuh: Finalize): mean path / variance or perturbation path
  float* g1;
  long long w_off[TT_LAYERS], b_off[TT_LAYERS];
};
// forward: x -> feat (+ the activation / eps images the backward pass re-reads); 5 launches
void tt_forward(const TtLane& ln, const TtStep& s, cudaStream_t st);
// backward: feat_grad -> g0 / g1 of the ten conv layers (atomic accumulation: the buffers must be zeroed); 3 launches
void tt_backward(const TtLane& ln, const TtStep& s, cudaStream_t st);
int tt_status();  // 0 ok; else the code of the first bounded mbarrier wait that timed out (synchronises)

}  // namespace brl
