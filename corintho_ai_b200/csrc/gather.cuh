// The one collective of the path behind the C ABI: an all-gather of the finished,
// un-augmented training samples over NCCL. The library does not link NCCL: the symbols are taken
// from the libnccl.so.2 that is already in the process (the host application's -- torch's bundled
// copy under Python), or loaded on first use. Only the stable C entry points are used; the few
// types are declared here as nccl.h declares them.
#ifndef CORINTHO_B200_GATHER_CUH
#define CORINTHO_B200_GATHER_CUH

#include <dlfcn.h>

#include "common.cuh"

namespace cb200 {

struct NcclUniqueId {
  char internal[128];
};
struct NcclApi {
  void *lib = nullptr;
  int (*GetUniqueId)(NcclUniqueId *) = nullptr;
  int (*CommInitRank)(void **comm, int nranks, NcclUniqueId id, int rank) = nullptr;
  int (*CommDestroy)(void *comm) = nullptr;
  int (*CommCount)(void *comm, int *count) = nullptr;
  int (*CommUserRank)(void *comm, int *rank) = nullptr;
  int (*AllGather)(const void *send, void *recv, size_t count, int dtype, void *comm, cudaStream_t st) = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
  bool ok = false;
};
constexpr int kNcclInt32 = 2, kNcclFloat32 = 7;  // ncclDataType_t

inline NcclApi &nccl_api() {
  static NcclApi a;
  if (a.lib) return a;
  a.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);  // the copy the host already uses
  if (!a.lib) a.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!a.lib) a.lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!a.lib) return a;
  a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(a.lib, "ncclGetUniqueId");
  a.CommInitRank = (decltype(a.CommInitRank))dlsym(a.lib, "ncclCommInitRank");
  a.CommDestroy = (decltype(a.CommDestroy))dlsym(a.lib, "ncclCommDestroy");
  a.CommCount = (decltype(a.CommCount))dlsym(a.lib, "ncclCommCount");
  a.CommUserRank = (decltype(a.CommUserRank))dlsym(a.lib, "ncclCommUserRank");
  a.AllGather = (decltype(a.AllGather))dlsym(a.lib, "ncclAllGather");
  a.GetErrorString = (decltype(a.GetErrorString))dlsym(a.lib, "ncclGetErrorString");
  a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.CommCount && a.CommUserRank && a.AllGather &&
         a.GetErrorString;
  return a;
}

#define CB_NCCL(expr)                                                                          \
  do {                                                                                         \
    const int _r = (expr);                                                                     \
    if (_r != 0)                                                                               \
      return cb200::set_error(CB200_ERR_CUDA, std::string(#expr) + ": " +                      \
                                                  cb200::nccl_api().GetErrorString(_r));       \
  } while (0)

// rows of rank r (padded block r of `padded`) -> their place in the contiguous result
__global__ void k_compact_gathered(const float *__restrict__ padded, float *__restrict__ out,
                                   const int32_t *__restrict__ counts, int world, int max_rows, int width) {
  const int r = blockIdx.y;
  int base = 0;
  for (int i = 0; i < r; ++i) base += counts[i];
  const size_t n = (size_t)counts[r] * width;
  const float *src = padded + (size_t)r * max_rows * width;
  float *dst = out + (size_t)base * width;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = src[i];
}

}  // namespace cb200
#endif
