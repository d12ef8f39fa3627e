// Self-play generation from a C++ host through the C ABI only (include/corintho_b200.h) -- what the
// reference's generation loop (corintho_ai/python/main.pyx:123-219: play_games + get_samples) becomes
// when the evaluation stays on the GPU. No Python, no torch.
//
//   g++ -O2 -std=c++17 -Iinclude examples/selfplay_main.cpp -Lcorintho_ai_b200 -lcorintho_b200
//       -Wl,-rpath,$PWD/corintho_ai_b200 -o examples/selfplay_main        (one command line)
//   examples/selfplay_main [games] [sims] [seed] [weights.f32 | -] [out_prefix | -]
//
// weights.f32: 127 997 little-endian floats in the layout of cb200_trainer_set_weights (BatchNorm
// folded; tflite_import.load_tflite_weights(...).tofile(...) writes it); "-" = a random-init
// network. out_prefix: writes <prefix>_game_states.f32, _evaluation_labels.f32,
// _probability_labels.f32 (the three arrays of main.pyx:202-204, rows in completion order) and
// <prefix>_game_of.i32.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>
#include <cmath>
#include <random>
#include <string>
#include <vector>

#include "corintho_b200.h"

static const size_t kWeights = 127997;

static std::vector<float> random_init(unsigned seed) {  // Glorot-uniform kernels, zero biases (wrapper.py:256-271)
  std::vector<float> w;
  w.reserve(kWeights);
  std::mt19937 gen(seed);
  const int dims[14] = {70, 100, 100, 100, 100, 100, 100, 100, 100, 100, 100, 100, 100, 97};
  // inference-time BatchNorm at its initial state is y / sqrt(1 + 1e-3): folded into the next Dense
  const float bn = 1.0f / std::sqrt(1.0f + 1e-3f);
  for (int l = 0; l < 13; ++l) {
    const int K = dims[l], N = dims[l + 1];
    // the two heads (1 and 96 outputs) are separate Dense layers in the reference
    for (int k = 0; k < K; ++k)
      for (int o = 0; o < N; ++o) {
        const int fan_out = l < 12 ? N : (o == 0 ? 1 : 96);
        const float lim = std::sqrt(6.0f / (float)(K + fan_out));
        std::uniform_real_distribution<float> u(-lim, lim);
        w.push_back(u(gen) * (l == 0 ? 1.0f : bn));
      }
    for (int o = 0; o < N; ++o) w.push_back(0.0f);
  }
  return w;
}

static bool dump(const std::string &path, const void *p, size_t bytes) {
  FILE *f = fopen(path.c_str(), "wb");
  if (!f) return false;
  const bool ok = fwrite(p, 1, bytes, f) == bytes;
  fclose(f);
  return ok;
}

int main(int argc, char **argv) {
  const int games = argc > 1 ? atoi(argv[1]) : 256;
  const int sims = argc > 2 ? atoi(argv[2]) : 200;
  const int seed = argc > 3 ? atoi(argv[3]) : 12345;
  const std::string wfile = argc > 4 ? argv[4] : "-";
  const std::string out = argc > 5 ? argv[5] : "-";
  if (cb200_device_count() < 1) {
    fprintf(stderr, "no CUDA device: the engine has no CPU fallback\n");
    return 2;
  }
  std::vector<float> w;
  if (wfile == "-") {
    w = random_init(0);
  } else {
    w.resize(kWeights);
    FILE *f = fopen(wfile.c_str(), "rb");
    if (!f || fread(w.data(), sizeof(float), kWeights, f) != kWeights) {
      fprintf(stderr, "cannot read %zu floats from %s\n", kWeights, wfile.c_str());
      return 2;
    }
    fclose(f);
  }
  // Trainer(num_games, log_folder, seed, max_searches, searches_per_eval, c_puct, epsilon, num_logged,
  //         num_threads, testing)  -- trainer.h:17-28, the values of toml/train.toml
  cb200_trainer *t = cb200_trainer_create(games, "", seed, sims, 16, 1.0f, 0.25f, 0, 1, 0);
  if (!t) {
    fprintf(stderr, "cb200_trainer_create: %s\n", cb200_last_error());
    return 1;
  }
  // precision 1 = bf16 tensor cores (random-init networks); 3 = bf16x3 for trained checkpoints
  int rc = cb200_trainer_set_weights(t, 0, w.data(), w.size(), wfile == "-" ? 1 : 3);
  if (rc == CB200_OK) rc = cb200_trainer_stream_samples(t, -1);
  const auto t0 = std::chrono::steady_clock::now();
  int done = 0;
  while (rc == CB200_OK && !done) {
    done = cb200_trainer_run_selfplay(t, 0, 0);  // until every game is over
    if (done < 0) rc = done;
  }
  const float *gs = nullptr, *ev = nullptr, *pr = nullptr;
  const int32_t *game_of = nullptr;
  int n = 0;
  if (rc == CB200_OK) rc = cb200_trainer_streamed_samples(t, &gs, &ev, &pr, &game_of, &n);
  const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  if (rc != CB200_OK) {
    fprintf(stderr, "error %d: %s\n", rc, cb200_last_error());
    cb200_trainer_destroy(t);
    return 1;
  }
  int64_t c[4] = {0, 0, 0, 0};
  cb200_trainer_counters(t, c);
  printf("games %d sims/move %d: %lld simulations, %lld moves, %d samples (x8 symmetries) in %.3f s = %.3e sims/s; "
         "score %.4f, mate length %.3f\n",
         games, sims, (long long)c[0], (long long)c[1], n, secs, (double)c[0] / secs, cb200_trainer_score(t),
         cb200_trainer_avg_mate_length(t));
  bool ok = c[1] == n;
  if (out != "-") {
    ok = ok && dump(out + "_game_states.f32", gs, (size_t)n * 8 * 70 * sizeof(float)) &&
         dump(out + "_evaluation_labels.f32", ev, (size_t)n * 8 * sizeof(float)) &&
         dump(out + "_probability_labels.f32", pr, (size_t)n * 8 * 96 * sizeof(float)) &&
         dump(out + "_game_of.i32", game_of, (size_t)n * sizeof(int32_t));
  }
  cb200_trainer_destroy(t);
  return ok ? 0 : 1;
}
