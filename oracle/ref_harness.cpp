// TEST INFRASTRUCTURE ONLY -- never linked into or called by the product path.
//
// C-ABI shim over the UNMODIFIED reference sources (compiled where they lie under
// /root/reference by oracle/Makefile; nothing from the reference is copied into this repo).
// It exposes the reference's public classes -- Game (cpp/include/game.h:19-135), Move codec
// (cpp/include/move.h:21-67), Trainer (cpp/include/trainer.h:17-53), Tourney
// (cpp/include/tourney.h:12-46) -- plus the rule tables of
// cpp/include/util.h as plain functions so that Python tests (ctypes) and bench.py's
// cpu_baseline / --impl reference legs can run the real reference side by side with the
// oracle restatement (oracle/corintho_oracle.cpp) and with the CUDA engine.
//
// Packed game state used across this repo ("cstate", 16 bytes, see include/corintho_b200.h):
//   w0 : the 64 board bits, bit index row*16 + col*4 + {0 base,1 column,2 capital,3 frozen}
//        (same index order as Game::board_, game.cpp:141-150)
//   w1 : byte i (i<6) = pieces_[i] (game.h:127-131), byte 6 = to_play_, byte 7 = 0

#include <cstdint>
#include <cstring>

#include <bitset>
#include <string>
#include <vector>

#include <omp.h>

#include "game.h"
#include "move.h"
#include "tourney.h"
#include "trainer.h"
#include "util.h"

namespace {

Game unpack(const uint64_t st[2]) {
  int32_t board[4 * kBoardSize];
  for (int i = 0; i < 64; ++i) board[i] = (st[0] >> i) & 1;
  int32_t pieces[6];
  for (int i = 0; i < 6; ++i) pieces[i] = (st[1] >> (8 * i)) & 0xff;
  int32_t to_play = (st[1] >> 48) & 0xff;
  return Game(board, to_play, pieces);
}

// Game's members are private; recover them through the public NN encoding
// (game.cpp:45-58): board bits verbatim, pieces rotated so the mover comes first.
void pack(const Game &g, int to_play, uint64_t st[2]) {
  float gs[kGameStateSize];
  g.writeGameState(gs);
  uint64_t w0 = 0;
  for (int i = 0; i < 64; ++i)
    if (gs[i] != 0.0f) w0 |= 1ull << i;
  uint64_t w1 = 0;
  for (int i = 0; i < 6; ++i) {
    int idx = (to_play * 3 + i) % 6;
    uint64_t cnt = static_cast<uint64_t>(gs[64 + i] * 4.0f + 0.5f);
    w1 |= cnt << (8 * idx);
  }
  w1 |= static_cast<uint64_t>(to_play) << 48;
  st[0] = w0;
  st[1] = w1;
}

inline void mask_words(const std::bitset<kNumMoves> &m, uint32_t out[3]) {
  out[0] = out[1] = out[2] = 0;
  for (int i = 0; i < kNumMoves; ++i)
    if (m[i]) out[i >> 5] |= 1u << (i & 31);
}

inline uint64_t splitmix(uint64_t &s) {
  uint64_t z = (s += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

}  // namespace

extern "C" {

// ---- constants & tables (util.h) -------------------------------------------------------
int ref_num_moves() { return kNumMoves; }
int ref_game_state_size() { return kGameStateSize; }
void ref_line_breaker(int idx, uint32_t out[3]) { mask_words(line_breakers[idx], out); }
float ref_gamma_sample(int i) { return gamma_samples[i]; }
int ref_space_symmetry(int k, int j) { return space_symmetries[k][j]; }
int ref_move_symmetry(int k, int j) { return move_symmetries[k][j]; }

// ---- move codec (move.cpp:11-42, 80-108) ------------------------------------------------
// out = {is_place, piece, row_from, col_from, row_to, col_to}
void ref_move_decode(int id, int out[6]) {
  Move m{id};
  out[0] = m.move_type() == Move::MoveType::kPlace;
  out[1] = out[0] ? m.piece_type() : -1;
  out[2] = m.row_from();
  out[3] = m.col_from();
  out[4] = m.row_to();
  out[5] = m.col_to();
}
int ref_encode_place(int row, int col, int piece) { return encodePlace(Space{row, col}, piece); }
int ref_encode_move(int r0, int c0, int r1, int c1) {
  return encodeMove(Space{r0, c0}, Space{r1, c1});
}

// ---- game rules ------------------------------------------------------------------------
void ref_game_start(uint64_t st[2]) {
  Game g;
  pack(g, 0, st);
}
int ref_game_legal(const uint64_t st[2], uint32_t mask[3]) {
  Game g = unpack(st);
  std::bitset<kNumMoves> legal;
  bool lines = g.getLegalMoves(legal);
  mask_words(legal, mask);
  return lines ? 1 : 0;
}
void ref_game_do_move(const uint64_t st[2], int move, uint64_t out[2]) {
  Game g = unpack(st);
  int to_play = (st[1] >> 48) & 0xff;
  g.doMove(move);
  pack(g, 1 - to_play, out);
}
void ref_game_encode(const uint64_t st[2], float out[70]) {
  Game g = unpack(st);
  g.writeGameState(out);
}

// One "game-logic step" per state (BASELINE.json configs[1]): legal mask + lines flag,
// terminal result (node.cpp:256-271: 0 none, 1 loss, 2 draw), NN encoding, and the state
// after the (rnd % n_legal)-th legal move in ascending id order (unchanged if terminal).
// flags[i] = result | is_lines<<2 | n_legal<<8 | chosen_move<<16 (chosen 0x7f if terminal).
void ref_game_step_batch(int64_t n, const uint64_t *states, const uint32_t *rnd,
                         uint32_t *masks, uint32_t *flags, uint64_t *next, float *enc,
                         int num_threads) {
  omp_set_num_threads(num_threads > 0 ? num_threads : 1);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    Game g = unpack(states + 2 * i);
    int to_play = (states[2 * i + 1] >> 48) & 0xff;
    std::bitset<kNumMoves> legal;
    bool lines = g.getLegalMoves(legal);
    mask_words(legal, masks + 3 * i);
    int n_legal = static_cast<int>(legal.count());
    int result = n_legal == 0 ? (lines ? kResultLoss : kResultDraw) : kResultNone;
    if (enc != nullptr) g.writeGameState(enc + kGameStateSize * i);
    int chosen = 0x7f;
    if (n_legal > 0) {
      int k = rnd[i] % n_legal;
      for (int m = 0; m < kNumMoves; ++m) {
        if (legal[m] && k-- == 0) {
          chosen = m;
          break;
        }
      }
      g.doMove(chosen);
      pack(g, 1 - to_play, next + 2 * i);
    } else {
      next[2 * i] = states[2 * i];
      next[2 * i + 1] = states[2 * i + 1];
    }
    flags[i] = result | (lines ? 4u : 0u) | (n_legal << 8) | (chosen << 16);
  }
}

// Reachable states from uniformly random legal play-outs of the reference rules; every
// position of every play-out (start and terminal positions included) is emitted until n
// states exist. Returns the number of play-outs used.
int64_t ref_gen_states(uint64_t seed, int64_t n, uint64_t *out) {
  uint64_t s = seed;
  int64_t count = 0, games = 0;
  while (count < n) {
    Game g;
    int to_play = 0;
    ++games;
    for (;;) {
      pack(g, to_play, out + 2 * count);
      if (++count >= n) break;
      std::bitset<kNumMoves> legal;
      g.getLegalMoves(legal);
      int n_legal = static_cast<int>(legal.count());
      if (n_legal == 0) break;
      int k = static_cast<int>(splitmix(s) % n_legal);
      int chosen = 0;
      for (int m = 0; m < kNumMoves; ++m)
        if (legal[m] && k-- == 0) {
          chosen = m;
          break;
        }
      g.doMove(chosen);
      to_play = 1 - to_play;
    }
  }
  return games;
}

// ---- Trainer (trainer.h:17-53) ---------------------------------------------------------
void *ref_trainer_create(int num_games, const char *log_folder, int seed, int max_searches,
                         int searches_per_eval, float c_puct, float epsilon, int num_logged,
                         int num_threads, int testing) {
  return new Trainer(num_games, std::string(log_folder ? log_folder : ""), seed, max_searches,
                     searches_per_eval, c_puct, epsilon, num_logged, num_threads, testing != 0);
}
void ref_trainer_destroy(void *h) { delete static_cast<Trainer *>(h); }
int ref_trainer_do_iteration(void *h, float *eval, float *probs, int to_play) {
  return static_cast<Trainer *>(h)->doIteration(eval, probs, to_play) ? 1 : 0;
}
int ref_trainer_num_requests(void *h, int to_play) {
  return static_cast<Trainer *>(h)->num_requests(to_play);
}
void ref_trainer_write_requests(void *h, float *game_states, int to_play) {
  static_cast<Trainer *>(h)->writeRequests(game_states, to_play);
}
int ref_trainer_num_samples(void *h) { return static_cast<Trainer *>(h)->num_samples(); }
void ref_trainer_write_samples(void *h, float *game_states, float *eval_samples,
                               float *prob_samples) {
  static_cast<Trainer *>(h)->writeSamples(game_states, eval_samples, prob_samples);
}
float ref_trainer_score(void *h) { return static_cast<Trainer *>(h)->score(); }
float ref_trainer_avg_mate_length(void *h) {
  return static_cast<Trainer *>(h)->avg_mate_length();
}
void ref_trainer_write_scores(void *h, const char *file) {
  static_cast<Trainer *>(h)->writeScores(std::string(file));
}


// ---- Tourney (tourney.h:12-46; Match behind it, match.h:33-101) -------------------------------
void *ref_tourney_create(int num_threads, const char *log_folder) {
  return new Tourney(num_threads, std::string(log_folder ? log_folder : ""));
}
void ref_tourney_destroy(void *h) { delete static_cast<Tourney *>(h); }
void ref_tourney_add_player(void *h, int player_id, int model_id, int max_searches,
                            int searches_per_eval, float c_puct, float epsilon, int random) {
  static_cast<Tourney *>(h)->addPlayer(player_id, model_id, max_searches, searches_per_eval, c_puct,
                                       epsilon, random != 0);
}
void ref_tourney_add_match(void *h, int player1, int player2, int logging) {
  static_cast<Tourney *>(h)->addMatch(player1, player2, logging != 0);
}
int ref_tourney_all_done(void *h) { return static_cast<Tourney *>(h)->all_done() ? 1 : 0; }
int ref_tourney_num_requests(void *h, int id) { return static_cast<Tourney *>(h)->num_requests(id); }
void ref_tourney_write_requests(void *h, float *game_states, int id) {
  static_cast<Tourney *>(h)->writeRequests(game_states, id);
}
void ref_tourney_do_iteration(void *h, float *eval, float *probs, int id) {
  static_cast<Tourney *>(h)->doIteration(eval, probs, id);
}
void ref_tourney_write_scores(void *h, const char *file) {
  static_cast<Tourney *>(h)->writeScores(std::string(file));
}

}  // extern "C"
