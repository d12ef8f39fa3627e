// K5 (fp32 parity variant): the policy/value network evaluated straight from packed cstates.
// Architecture = corintho_ai/python/wrapper.py:256-271 with BatchNorm folded into the next
// Dense by the caller: 70 -> 12 x [Dense 100, ReLU] -> {tanh value, softmax 96 policy}.
// Replaces the Keras predict call of the reference loop (corintho_ai/python/main.pyx:70-83).
//
// One CTA = 128 positions through all 13 layers; activations stay in shared memory
// (k-major, so a lane reads 4 positions with one 128-bit load), each layer's weights are
// staged once per CTA. fp32 FFMA throughout -- this is the 1e-5 parity evaluator; the
// tensor-core bf16 variant lives in mlp_tc.cuh.
#ifndef CORINTHO_B200_MLP_CUH
#define CORINTHO_B200_MLP_CUH

#include "common.cuh"

namespace cb200 {

constexpr int kNetLayers = 13;    // 12 hidden + fused head
constexpr int kNetHidden = 100;
constexpr int kNetHead = 97;      // value + 96 policy logits
constexpr int kNetNPad = 104;     // 8 warps x 13 outputs
constexpr int kNetTile = 128;     // positions per CTA
constexpr int kNetThreads = 256;
constexpr size_t kNetWeightFloats = 127997;

// fp32 weights in kernel layout: per layer [K_l][kNetNPad] then bias[kNetNPad]
struct NetF32 {
  float *w = nullptr;  // device
  size_t layer_off[kNetLayers];
  int K[kNetLayers], N[kNetLayers];
  bool ready = false;
};

constexpr size_t kMlpSmemBytes =
    (size_t)(2 * kNetHidden * kNetTile + kNetHidden * kNetNPad + kNetNPad + 2 * kNetTile) * 4;

__global__ void __launch_bounds__(kNetThreads, 1)
    k_mlp_f32(const float *__restrict__ W, const ulonglong2 *__restrict__ states,
              const int32_t *__restrict__ n_ptr, int n_static, float *__restrict__ eval,
              float *__restrict__ probs, int32_t *__restrict__ zero2) {
  extern __shared__ float smf[];
  float *act0 = smf;                              // [100][128]
  float *act1 = act0 + kNetHidden * kNetTile;     // [100][128]
  float *wsm = act1 + kNetHidden * kNetTile;      // [100][104]
  float *bsm = wsm + kNetHidden * kNetNPad;       // [104]
  float *rmax = bsm + kNetNPad;                   // [128]
  float *rsum = rmax + kNetTile;                  // [128]
  const int n = n_ptr ? *n_ptr : n_static;
  if (zero2 && blockIdx.x == 0 && threadIdx.x == 0) zero2[0] = 0, zero2[2] = 0;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  for (int tile = blockIdx.x; tile * kNetTile < n; tile += gridDim.x) {
    const int p0 = tile * kNetTile;
    // input encoding (game.cpp:45-58) straight from the packed state
    for (int idx = t; idx < CB200_STATE_SIZE * kNetTile; idx += kNetThreads) {
      const int k = idx / kNetTile, p = idx - k * kNetTile;
      float v = 0.0f;
      if (p0 + p < n) {
        const ulonglong2 s = states[p0 + p];
        v = encode_elem(CState{s.x, s.y}, k);
      }
      act0[k * kNetTile + p] = v;
    }
    float *ain = act0, *aout = act1;
    size_t woff = 0;
    for (int layer = 0; layer < kNetLayers; ++layer) {
      const int K = layer == 0 ? CB200_STATE_SIZE : kNetHidden;
      __syncthreads();
      for (int idx = t; idx < K * kNetNPad; idx += kNetThreads) wsm[idx] = W[woff + idx];
      if (t < kNetNPad) bsm[t] = W[woff + (size_t)K * kNetNPad + t];
      woff += (size_t)(K + 1) * kNetNPad;
      __syncthreads();
      float acc[4][13];
#pragma unroll
      for (int j = 0; j < 13; ++j) {
        const float b = bsm[warp * 13 + j];
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i][j] = b;
      }
      for (int k = 0; k < K; ++k) {
        const float4 a = *reinterpret_cast<const float4 *>(ain + k * kNetTile + lane * 4);
        const float *wr = wsm + k * kNetNPad + warp * 13;
#pragma unroll
        for (int j = 0; j < 13; ++j) {
          const float w = wr[j];
          acc[0][j] = fmaf(a.x, w, acc[0][j]);
          acc[1][j] = fmaf(a.y, w, acc[1][j]);
          acc[2][j] = fmaf(a.z, w, acc[2][j]);
          acc[3][j] = fmaf(a.w, w, acc[3][j]);
        }
      }
      const bool last = layer == kNetLayers - 1;
#pragma unroll
      for (int j = 0; j < 13; ++j) {
        const int o = warp * 13 + j;
        if (o < kNetHidden) {
          float4 v = make_float4(acc[0][j], acc[1][j], acc[2][j], acc[3][j]);
          if (!last) v.x = fmaxf(v.x, 0.f), v.y = fmaxf(v.y, 0.f), v.z = fmaxf(v.z, 0.f), v.w = fmaxf(v.w, 0.f);
          *reinterpret_cast<float4 *>(aout + o * kNetTile + lane * 4) = v;
        }
      }
      float *tmp = ain;
      ain = aout, aout = tmp;
    }
    __syncthreads();
    // heads: ain[0][p] = value pre-activation, ain[1..96][p] = policy logits
    if (t < kNetTile) {
      float mx = -INFINITY;
      for (int j = 1; j < kNetHead; ++j) mx = fmaxf(mx, ain[j * kNetTile + t]);
      float s = 0.0f;
      for (int j = 1; j < kNetHead; ++j) s += expf(ain[j * kNetTile + t] - mx);
      rmax[t] = mx, rsum[t] = s;
      if (p0 + t < n) eval[p0 + t] = tanhf(ain[t]);
    }
    __syncthreads();
    const int cnt = min(kNetTile, n - p0);
    for (int idx = t; idx < cnt * CB200_NUM_MOVES; idx += kNetThreads) {
      const int p = idx / CB200_NUM_MOVES, j = idx - p * CB200_NUM_MOVES;
      probs[(size_t)p0 * CB200_NUM_MOVES + idx] = expf(ain[(j + 1) * kNetTile + p] - rmax[p]) / rsum[p];
    }
    __syncthreads();
  }
}

// Re-layout the C-ABI weight vector (include/corintho_b200.h cb200_trainer_set_weights) into
// the kernel layout and upload it.
inline int net_f32_upload(NetF32 &net, const float *weights) {
  std::string &err = last_error_ref();
  (void)err;
  size_t total = 0;
  for (int l = 0; l < kNetLayers; ++l) {
    net.K[l] = l == 0 ? CB200_STATE_SIZE : kNetHidden;
    net.N[l] = l == kNetLayers - 1 ? kNetHead : kNetHidden;
    net.layer_off[l] = total;
    total += (size_t)(net.K[l] + 1) * kNetNPad;
  }
  float *host = (float *)calloc(total, sizeof(float));
  if (!host) return set_error(CB200_ERR_ARG, "out of host memory");
  const float *src = weights;
  for (int l = 0; l < kNetLayers; ++l) {
    float *dstw = host + net.layer_off[l];
    for (int k = 0; k < net.K[l]; ++k)
      for (int o = 0; o < net.N[l]; ++o) dstw[(size_t)k * kNetNPad + o] = src[(size_t)k * net.N[l] + o];
    src += (size_t)net.K[l] * net.N[l];
    float *dstb = dstw + (size_t)net.K[l] * kNetNPad;
    for (int o = 0; o < net.N[l]; ++o) dstb[o] = src[o];
    src += net.N[l];
  }
  if (!net.w) {
    cudaError_t e = cudaMalloc(&net.w, total * sizeof(float));
    if (e != cudaSuccess) {
      free(host);
      return set_error(CB200_ERR_CUDA, std::string("cudaMalloc net: ") + cudaGetErrorString(e));
    }
  }
  cudaError_t e = cudaMemcpy(net.w, host, total * sizeof(float), cudaMemcpyHostToDevice);
  free(host);
  if (e != cudaSuccess) return set_error(CB200_ERR_CUDA, cudaGetErrorString(e));
  net.ready = true;
  return CB200_OK;
}

inline int launch_mlp_f32(const NetF32 &net, const ulonglong2 *d_states, const int32_t *d_n,
                          int n_static, int n_max, float *d_eval, float *d_probs,
                          int32_t *zero2 = nullptr, cudaStream_t stream = nullptr,
                          bool use_stream = false) {
  static bool attr_set[16] = {false};
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (dev < 16 && !attr_set[dev]) {
    CB_CUDA(cudaFuncSetAttribute(k_mlp_f32, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)kMlpSmemBytes));
    attr_set[dev] = true;
  }
  if (n_max <= 0) return CB200_OK;
  int tiles = (n_max + kNetTile - 1) / kNetTile;
  int grid = tiles < sms ? tiles : sms;
  k_mlp_f32<<<grid, kNetThreads, kMlpSmemBytes, use_stream ? stream : cur_stream()>>>(
      net.w, d_states, d_n, n_static, d_eval, d_probs, zero2);
  CB_LAUNCHED();
  CB_CUDA(cudaGetLastError());
  return CB200_OK;
}

}  // namespace cb200
#endif
