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
  float* part;           // weight-gradient partials of the backward CTA groups: [group][layer blocks] (tt_reduce_kernel sums them)
  // fc layer operands, per 128-window M-tile [300 k-chunks][128 windows][16 B], k' = t * 80 + c:
  unsigned char* fimg;   // fp16 module-2 output (features)
  unsigned char* fbimg;  // the same in bf16 (A operand of the weight-gradient GEMM)
  unsigned char* f2img;  // second operand of the forward GEMM: bf16 f^2 (LRT) or fp16 f * s_in (Flipout)
  unsigned char* f2bimg; // Flipout: bf16 f * s_in (A operand of the perturbation-path weight gradient)
  unsigned char* gfimg;  // bf16 gradient w.r.t. the features (written by the fc input-gradient kernel)
  unsigned char* fcblob; // fc weight images [300][64][8]: fp16 mu | bf16 mu | bf16 sigma^2 or (W - mu) | fp16 of the same
};
size_t tt_lane_bytes(long long B);
void tt_carve(unsigned char* base, long long B, TtLane& ln);

struct TtStep {
  const float* x;        // [B,30,18]
  long long B;
  int mode;              // BRL_MODE_LRT / BRL_MODE_FLIPOUT (dual contraction) or BRL_MODE_DET / BRL_MODE_WS (one contraction with the
                         // weights `mu`: the HNN / MC-dropout step, the weight-sampling ELBO of the radial guide)
  const float* mu;       // [P]
  const float* sigma;    // [P]  (LRT)
  const float* wsamp;    // [P]  (Flipout: the particle's weight draw)
  NoiseRef eps[TT_LAYERS];          // LRT eps streams (injected tensor [B, N*30] or Philox)
  const float* sgn_in[TT_LAYERS];   // Flipout [B, Cin]
  const float* sgn_out[TT_LAYERS];  // Flipout [B, Cout]
  NoiseRef drop[TT_LAYERS];         // single-contraction modes: dropout site behind each conv layer (masks [B, N, 30] or Philox)
  float keep[TT_LAYERS];            // its keep probability (1 = no site / dropout off)
  const float* sgn_fc_in;  // Flipout: s_in of the fc layer [B, 2400]
  long long w_off_fc, b_off_fc;
  float* g0;             // flat gradient accumulators [P] (brl_kernels.cuh: Finalize): mean path / variance or perturbation path
  float* g1;
  long long w_off[TT_LAYERS], b_off[TT_LAYERS];
};
// forward of the ten conv layers: x -> feature images (+ the activation / eps images the backward pass re-reads); 6 launches
// side stream + two events of the calling lane: independent kernels of a step (weight packing next to the window packing, the fc
// weight gradient next to the conv backward chain) fork to `side` and join back; all of it is stream-capturable
struct TtSide { cudaStream_t side; cudaEvent_t fork, join; };
void tt_forward(const TtLane& ln, const TtStep& s, cudaStream_t st, const TtSide& sd);
// fc layer GEMMs on the tensor pipe.  Forward: both contractions as split-K partial sums into `part` ([2][B][64] fp32, the
// split-K scratch layout of brl_gemm.cu) -- the per-layer engine's split-K epilogue then applies bias / eps * sqrt(var) / signs /
// ReLU.  Backward: dpre / dsec = compact [B, 64] gradients (bwd_act_kernel) -> feature-gradient image + g0 / g1 of the fc layer.
void tt_fc_forward(const TtLane& ln, const TtStep& s, float* part, cudaStream_t st);
void tt_fc_backward(const TtLane& ln, const TtStep& s, const float* dpre, const float* dsec, cudaStream_t st, const TtSide& sd);
// Everything between the fc GEMM and the fc layer's backward GEMMs in ONE launch: fc epilogue over `part`, head forward, softplus /
// threshold, ELBO likelihood (+ its sums in acc[0..1]), head backward (its weight gradients into g0 / g1) and the fc layer's
// activation backward (dpre / dsec, compact [B, 64]).
struct TtTail {
  const float* part;
  const float* y;
  float gscale;           // c_nll / particles (launch_nll_elbo)
  int compute_grads;
  double* acc;
  float* out;             // [B,2]
  float *dpre, *dsec;
  NoiseRef eps_fc, eps_head;
  const float *sout_fc, *sin_head, *sout_head;
  long long hw_off, hb_off;  // flat offsets of the head layer's weight / bias
  int loss_kind = 0;         // single-contraction modes: 0 ELBO likelihood, 1 F.gaussian_nll_loss (HNN step)
  float keep_fc = 1.0f;      // dropout site behind the fc layer
  NoiseRef drop_fc{};
};
void tt_tail(const TtStep& s, const TtTail& t, cudaStream_t st);
// backward: feature-gradient image -> g0 / g1 of the ten conv layers (accumulated: the buffers must be zeroed); 4 launches
void tt_backward(const TtLane& ln, const TtStep& s, cudaStream_t st, const TtSide& sd);
// debug: device buffer int64[6 launches][4 layers][16] receiving clock64 stamps of CTA (0, layer) of the three forward and three
// backward level launches (nullptr = off)
void tt_trace(long long* device_buf);
int tt_status();  // 0 ok; else the code of the first bounded mbarrier wait that timed out (synchronises)

}  // namespace brl
