// K5 (bf16 tensor-core variant) -- placeholder until the tcgen05 kernel lands (next commit).
#ifndef CORINTHO_B200_MLP_TC_CUH
#define CORINTHO_B200_MLP_TC_CUH
#include "common.cuh"
namespace cb200 {
struct NetTC {
  void *w = nullptr;
  bool ready = false;
};
inline int net_tc_upload(NetTC &, const float *) {
  return set_error(CB200_ERR_STATE, "bf16 tcgen05 evaluator not built yet");
}
inline void net_tc_free(NetTC &net) {
  if (net.w) cudaFree(net.w);
  net.w = nullptr;
}
inline int launch_mlp_tc(const NetTC &, const ulonglong2 *, const int32_t *, int, int, float *, float *) {
  return set_error(CB200_ERR_STATE, "bf16 tcgen05 evaluator not built yet");
}
}  // namespace cb200
#endif
