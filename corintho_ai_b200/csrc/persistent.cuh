// K7: persistent fused self-play for the phase with few live games.
//
// While thousands of games are live the lock-step loop [k_mlp_tc -> k_iterate] keeps every SM
// busy, but each launch lasts as long as its slowest game and every iteration pays two kernel
// boundaries. Once all live games fit on the GPU at 8 games per SM, one CTA per SM takes 8 games
// and loops by itself:
//     game step of its 8 games (one warp each: run_game = SelfPlayer::doIteration)
//  -> the tensor-core network on the <= 128 leaf positions they queued (tc_forward, one tile)
// without leaving the SM, so a CTA only ever waits for its own 8 games. Requests and answers
// travel through CTA-private rows of small global buffers (L2 resident). Per-game order of
// operations is unchanged, hence so is every result (tests compare against the lock-step run).
#ifndef CORINTHO_B200_PERSISTENT_CUH
#define CORINTHO_B200_PERSISTENT_CUH

#include "mlp_tc.cuh"
#include "tree.cuh"

namespace cb200 {

constexpr int kPsWarps = 8;                  // games per CTA
constexpr int kPsRows = 128;                 // request rows per CTA = one network tile
constexpr size_t kPsTreeSmemOff = (kTcSmemBytes + 127) / 128 * 128;
constexpr size_t kPsSmemBytes = kPsTreeSmemOff + kPsWarps * sizeof(WarpSm);

// live games in ascending index order (single thread: a few thousand flags, once per switch)
__global__ void k_live_list(TreeParams P, int32_t *__restrict__ list, int32_t *__restrict__ count) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  int n = 0;
  for (int g = 0; g < P.num_games; ++g)
    if (!P.ctl[(size_t)g * kCtlWords + CW_DONE]) list[n++] = g;
  *count = n;
}

// out[0] += games still live when the CTA stopped, out[1] = min error code, out[2] = max rounds
template <bool kFp16>
__global__ void __launch_bounds__(kTcThreads, 1)
    k_selfplay_persistent(TreeParams P, const uint8_t *__restrict__ W,
                          const int32_t *__restrict__ game_list, int n_list,
                          const float *eval0, const float *probs0, long pcs0, int first_external,
                          float *eval, float *probs, int ld, ulonglong2 *packed, int max_rounds,
                          int iteration0, int32_t *__restrict__ out) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ int32_t s_ctr[4];  // [0] requests of this round, [1] live games, [2] error
  TcState S;
  tc_setup(S, smem);
  WarpSm *sm_all = reinterpret_cast<WarpSm *>(smem + kPsTreeSmemOff);
  const int warp = threadIdx.x >> 5;
  const int slot = blockIdx.x * kPsWarps + warp;
  const int g = slot < n_list ? game_list[slot] : -1;
  const int row0 = blockIdx.x * kPsRows;
  if (threadIdx.x == 0) s_ctr[2] = 0;  // published by the first barrier of the loop
  int round = 0, live = 0;
  for (; round < max_rounds; ++round) {
    if (threadIdx.x == 0) s_ctr[0] = 0, s_ctr[1] = 0;
    __syncthreads();
    if (g >= 0) {
      const bool ext = first_external && round == 0;
      run_game<true>(P, g, sm_all[warp], ext ? eval0 : eval, ext ? probs0 : probs, 1,
                     ext ? pcs0 : (long)ld, nullptr, -1, iteration0 + round, 0, &s_ctr[0],
                     &s_ctr[1], &s_ctr[2], row0, packed);
    }
    __syncthreads();
    const int n = s_ctr[0];
    live = s_ctr[1];
    __syncthreads();  // everybody has read the counters before thread 0 clears them again
    // every round ends with the network, so the answers of all queued requests are in the
    // CTA's rows whenever the kernel stops (a later launch continues from there)
    if (n > 0) tc_forward<kFp16>(S, W, packed + row0, n, 0, eval + row0, probs + row0, ld);
    if (live == 0) {
      ++round;
      break;
    }
  }
  tc_teardown(S);
  if (threadIdx.x == 0) {
    if (live) atomicAdd(out, live);
    if (s_ctr[2]) atomicMin(out + 1, s_ctr[2]);
    atomicMax(out + 2, round);
  }
}

}  // namespace cb200
#endif
