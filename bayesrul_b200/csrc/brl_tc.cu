#include "brl_tc.cuh"
namespace brl {
struct TcState { int net; };
TcState* tc_create(int net) { return new TcState{net}; }
void tc_destroy(TcState* s) { delete s; }
bool tc_available(const TcState*) { return false; }
size_t tc_workspace_bytes(const TcState*, long long, long long) { return 0; }
const char* tc_forward(TcState*, const float*, long long, long long, const float*, long long, float, const brl_noise*, float*,
                       void*, size_t, cudaStream_t) {
  return "bayesrul_b200: tensor-core engine not available for this net";
}
}  // namespace brl
