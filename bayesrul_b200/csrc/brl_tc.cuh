// tcgen05 / TMEM engine (fp16 operands, fp32 accumulate) -- interface used by brl_api.cu.
#pragma once
#include <cstddef>
#include <cuda_runtime.h>

struct brl_noise;

namespace brl {
struct TcState;
TcState* tc_create(int net);
void tc_destroy(TcState*);
bool tc_available(const TcState*);
int tc_status(const TcState*);  // 0 ok; else the code of the first mbarrier wait that timed out (synchronises)
void tc_timing(TcState*, bool enable);
void tc_timing_read(TcState*, double ms[2], long long launches[2]);  // [conv, fc]; synchronises the recorded events
void tc_trace(TcState*, long long* device_buf);
size_t tc_workspace_bytes(const TcState*, long long B, long long S);
// returns nullptr on success, else a static error string.  pack_x = false re-uses the fp16 window images a previous
// call built in the same workspace for the same x (later chunks of MC samples of one batch).
const char* tc_forward(TcState*, const float* x, long long B, long long S, const float* weights, long long w_sample_stride,
                       float p_dropout, const brl_noise* noise, float* out, void* ws, size_t ws_bytes, bool pack_x,
                       cudaStream_t st);
}  // namespace brl
