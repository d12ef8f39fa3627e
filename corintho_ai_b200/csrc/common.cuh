// Shared host/device plumbing of libcorintho_b200.so (single translation unit: engine.cu).
#ifndef CORINTHO_B200_COMMON_CUH
#define CORINTHO_B200_COMMON_CUH

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>
#include <string>

#include "../../include/corintho_b200.h"
#include "corintho_tables.h"
#include "rules.cuh"

namespace cb200 {

// ---- error plumbing (never throw across the C ABI) ------------------------------------------
inline std::string &last_error_ref() {
  static thread_local std::string s;
  return s;
}
inline int set_error(int code, const std::string &msg) {
  last_error_ref() = msg;
  return code;
}
#define CB_CUDA(expr)                                                                       \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) {                                                                \
      return cb200::set_error(CB200_ERR_CUDA, std::string(#expr) + ": " +                   \
                                                  cudaGetErrorString(_e));                  \
    }                                                                                       \
  } while (0)

constexpr int kMaxDevices = 16;
struct Globals {
  // launch stream per device (cb200_set_stream applies to the device that is current when it is
  // called); devices past kMaxDevices share the last entry
  cudaStream_t stream_of[kMaxDevices] = {};
  std::atomic<int64_t> launches{0};
  bool tables_ready[kMaxDevices] = {false};
};
inline Globals &G() {
  static Globals g;
  return g;
}
inline int cur_device_slot() {
  int dev = 0;
  cudaGetDevice(&dev);
  return dev < 0 ? 0 : (dev < kMaxDevices ? dev : kMaxDevices - 1);
}
inline cudaStream_t cur_stream() { return G().stream_of[cur_device_slot()]; }
inline void set_cur_stream(cudaStream_t s) { G().stream_of[cur_device_slot()] = s; }
#define CB_LAUNCHED() (cb200::G().launches.fetch_add(1, std::memory_order_relaxed))

// ---- device-resident rule data (uploaded once per device by ensure_tables) ------------------
// line-breaker masks padded to 4 words so one 128-bit load fetches a mask
__device__ __align__(16) uint32_t d_line_breakers[103 * 4];  // entry 102 = all ones (no line)
__device__ float d_gamma[1024];
__device__ uint32_t d_inv32[128];  // floor(2^32 / d) for d = 1..127 ([0] unused; [1] = 2^32 - 1)
// K1 fast path (rules.cuh: build_move_lut / build_nth_lut); copied to shared memory by the kernel
__device__ __align__(16) uint32_t d_move_lut[96 * kMoveLutWords];
__device__ __align__(16) uint8_t d_nth_lut[256 * 8];

struct DeviceLB {
  __device__ __forceinline__ const uint32_t *operator()(int idx) const {
    return d_line_breakers + 4 * idx;
  }
};

inline int ensure_tables() {
  int dev = 0;
  CB_CUDA(cudaGetDevice(&dev));
  if (dev < 16 && G().tables_ready[dev]) return CB200_OK;
  static uint32_t lb[103 * 4];
  lb[408] = lb[409] = lb[410] = 0xFFFFFFFFu, lb[411] = 0;
  for (int i = 0; i < 102; ++i) {
    lb[4 * i] = kCLineBreakers[i][0], lb[4 * i + 1] = kCLineBreakers[i][1];
    lb[4 * i + 2] = kCLineBreakers[i][2], lb[4 * i + 3] = 0;
  }
  CB_CUDA(cudaMemcpyToSymbol(d_line_breakers, lb, sizeof(lb)));
  CB_CUDA(cudaMemcpyToSymbol(d_gamma, kCGammaBits, sizeof(kCGammaBits)));
  static uint32_t inv[128];
  inv[0] = 0, inv[1] = 0xFFFFFFFFu;
  for (uint32_t d = 2; d < 128; ++d) inv[d] = (uint32_t)(0x100000000ull / d);
  CB_CUDA(cudaMemcpyToSymbol(d_inv32, inv, sizeof(inv)));
  static uint32_t mv[96 * kMoveLutWords];
  static uint8_t nth[256 * 8];
  build_move_lut(mv), build_nth_lut(nth);
  CB_CUDA(cudaMemcpyToSymbol(d_move_lut, mv, sizeof(mv)));
  CB_CUDA(cudaMemcpyToSymbol(d_nth_lut, nth, sizeof(nth)));
  if (dev < 16) G().tables_ready[dev] = true;
  return CB200_OK;
}

constexpr unsigned kFull = 0xffffffffu;

}  // namespace cb200
#endif
