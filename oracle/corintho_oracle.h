/* TEST INFRASTRUCTURE ONLY -- the CPU oracle for the Corintho self-play path.
 *
 * A from-scratch CPU restatement (oracle/corintho_oracle.cpp) of the reference algorithm:
 *   game rules   corintho_ai/cpp/src/game.cpp:28-405, move.cpp:11-108, node.cpp:256-283
 *   tree search  corintho_ai/cpp/src/trainmc.cpp:110-696
 *   one game     corintho_ai/cpp/src/selfplayer.cpp:19-291
 *   game batch   corintho_ai/cpp/src/trainer.cpp:18-256
 * PARITY PINNED: tests/test_oracle_vs_ref.py checks it bit-for-bit against the compiled
 * reference (oracle/_ref/libcorintho_ref.so) and tests/test_golden.py against the committed
 * fixtures in tests/golden/ that were generated from the compiled reference.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library. The product (libcorintho_b200.so) never links or calls it.
 *
 * Packed game state ("cstate", 2 x u64): see include/corintho_b200.h.
 */
#ifndef CORINTHO_ORACLE_H
#define CORINTHO_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

void orc_line_breaker(int idx, uint32_t out[3]);
float orc_gamma_sample(int i);

void orc_move_decode(int id, int out[6]);
int orc_encode_place(int row, int col, int piece);
int orc_encode_move(int r0, int c0, int r1, int c1);

void orc_game_start(uint64_t st[2]);
int orc_game_legal(const uint64_t st[2], uint32_t mask[3]);
void orc_game_do_move(const uint64_t st[2], int move, uint64_t out[2]);
void orc_game_encode(const uint64_t st[2], float out[70]);
void orc_game_step_batch(int64_t n, const uint64_t *states, const uint32_t *rnd, uint32_t *masks,
                         uint32_t *flags, uint64_t *next, float *enc, int num_threads);

void *orc_trainer_create(int num_games, const char *log_folder, int seed, int max_searches,
                         int searches_per_eval, float c_puct, float epsilon, int num_logged,
                         int num_threads, int testing);
/* shard [first_game, first_game+num_games) of the same seed stream (engine multi-GPU sharding) */
void *orc_trainer_create_shard(int first_game, int num_games, const char *log_folder, int seed,
                               int max_searches, int searches_per_eval, float c_puct, float epsilon,
                               int num_logged, int num_threads, int testing);
void orc_trainer_destroy(void *h);
int orc_trainer_do_iteration(void *h, const float *eval, const float *probs, int to_play);
int orc_trainer_num_requests(void *h, int to_play);
void orc_trainer_write_requests(void *h, float *game_states, int to_play);
int orc_trainer_num_samples(void *h);
void orc_trainer_write_samples(void *h, float *game_states, float *eval_samples,
                               float *prob_samples);
float orc_trainer_score(void *h);
float orc_trainer_avg_mate_length(void *h);
/* exact counters for the metric: simulations = sum over moves of the mover's
 * searches_done_ at chooseMove time; moves = chooseMove calls; leaf_evals = requests served */
void orc_trainer_counters(void *h, int64_t out[3]);
/* per game: 0 unfinished, 1 first player lost, 2 draw, 3 first player won (util.h:58-61) */
void orc_trainer_game_results(void *h, int32_t *results);
/* white-box dump of one tree for engine-vs-oracle debugging:
 * out = {has_root, used_words, root_visits, root_result, root_all_visited, searches_done,
 *        root_eval_bits}; returns used_words and copies min(cap,used) arena words */
int orc_trainer_dump_tree(void *h, int game, int player, int64_t out[8], uint32_t *words,
                          int cap);


/* Tourney (corintho_ai/cpp/include/tourney.h:12-46) over Match (match.h:33-101): per-player
 * search budgets, random players, requests batched per model id. do_iteration returns -1 on an
 * arena/path overflow of any match, else 0. */
void *orc_tourney_create(int num_threads, const char *log_folder);
void orc_tourney_destroy(void *h);
void orc_tourney_add_player(void *h, int player_id, int model_id, int max_searches,
                            int searches_per_eval, float c_puct, float epsilon, int random);
void orc_tourney_add_match(void *h, int player1, int player2, int logging);
int orc_tourney_all_done(void *h);
int orc_tourney_num_requests(void *h, int id);
void orc_tourney_write_requests(void *h, float *game_states, int id);
int orc_tourney_do_iteration(void *h, const float *eval, const float *probs, int id);
void orc_tourney_write_scores(void *h, const char *file);

#ifdef __cplusplus
}
#endif
#endif
