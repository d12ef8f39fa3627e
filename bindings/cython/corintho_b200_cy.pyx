# distutils: language = c++
# cython: language_level=3
"""The reference-side binding of the B200 engine, in the reference's own binding language.

The two `cdef cppclass` blocks below are the reference's (corintho_ai/python/main.pyx:17-38,
corintho_ai/rating/tourney.pyx:15-32) with one change each: the file they are read from.
`play_games` / `run_tourney` restate the loops of main.pyx:123-219 and tourney.pyx:63-205 with the
evaluators passed in as callables (the reference calls Keras / tflite_runtime there, neither of
which exists in this image). Built by bindings/cython/setup.py; tests/test_cython_binding.py
compiles it and drives the engine through it."""
from libcpp cimport bool
from libcpp.string cimport string

import numpy as np

cdef extern from "corintho_b200.hpp":
    cdef cppclass Trainer:
        Trainer(
            int num_games,
            string log_folder,
            int seed,
            int max_searches,
            int searches_per_eval,
            float c_puct,
            float epsilon,
            int num_logged,
            int num_threads,
            bool testing,
        ) except +
        int num_requests(int to_play) except +
        int num_samples() except +
        float score() except +
        float avg_mate_length() except +
        void writeRequests(float *game_states, int to_play) except +
        void writeSamples(float *game_states, float *eval_samples, float *prob_samples) except +
        void writeScores(string file) except +
        bool doIteration(float *evaluations, float *probabilities, int to_play) except +
        # not in the reference: the network stays on the device (INTEGRATION.md section 3)
        void setWeights(int model, const float *weights, size_t n_floats, int precision) except +
        bool runSelfplay(int max_iterations, int stagger) except +

    cdef cppclass Tourney:
        Tourney(int num_threads, string log_folder) except +
        bool all_done() except +
        int num_requests(int id) except +
        void writeScores(string filename) except +
        void writeRequests(float *game_states, int id) except +
        void doIteration(float *eval, float *probs, int id) except +
        void addPlayer(
            int player_id,
            int model_id,
            int max_searches,
            int searches_per_eval,
            float c_puct,
            float epsilon,
            bool random) except +
        void addMatch(int player1, int player2, bool logging) except +
        int max_rows()          # not in the reference: the buffer bound its doIteration implies

cdef int _NUM_MOVES = 96
cdef int _GAME_STATE_SIZE = 70
cdef int _SYMMETRY_NUM = 8


def play_games(int num_games, str log_folder, int seed, int max_searches, int searches_per_eval,
               float c_puct, float epsilon, int num_logged, int num_threads, bint testing, evaluator,
               scores_file=None):
    """main.pyx:123-219 (play_games + get_samples): `evaluator(rows[n,70]) -> (evaluations[n],
    probabilities[n,96])` stands where the reference calls model.predict; in testing mode it is
    called as evaluator(rows, to_play) (main.pyx:172-182 picks one of two models).
    Returns (game_states[N*8,70], evaluation_labels[N*8], probability_labels[N*8,96], score,
    avg_mate_length)."""
    cdef Trainer *trainer = new Trainer(num_games, log_folder.encode(), seed, max_searches, searches_per_eval,
                                        c_puct, epsilon, num_logged, num_threads, testing)
    cdef int cap = num_games * searches_per_eval
    cdef float[::1] evaluations = np.zeros(cap, dtype=np.float32)
    cdef float[:, ::1] probabilities = np.zeros((cap, _NUM_MOVES), dtype=np.float32)
    cdef float[:, ::1] game_states = np.zeros((cap, _GAME_STATE_SIZE), dtype=np.float32)
    cdef int to_play = 0 if testing else -1
    cdef int n
    cdef float[:, ::1] sample_states
    cdef float[::1] evaluation_labels
    cdef float[:, ::1] probability_labels
    try:
        while True:
            if trainer.doIteration(&evaluations[0], &probabilities[0, 0], to_play):
                break
            n = trainer.num_requests(to_play)
            if n == 0:
                if testing:                  # main.pyx:151-154
                    to_play = 1 - to_play
                    continue
                raise RuntimeError("No requests during training")   # main.pyx:161-163
            trainer.writeRequests(&game_states[0, 0], to_play)
            rows = np.asarray(game_states)[:n]
            ev, pr = evaluator(rows, to_play) if testing else evaluator(rows)
            np.asarray(evaluations)[:n] = np.asarray(ev, dtype=np.float32).reshape(-1)
            np.asarray(probabilities)[:n] = pr
        if scores_file is not None:
            trainer.writeScores(str(scores_file).encode())
        n = trainer.num_samples()            # main.pyx:189-204
        sample_states = np.zeros((max(n, 1) * _SYMMETRY_NUM, _GAME_STATE_SIZE), dtype=np.float32)
        evaluation_labels = np.zeros(max(n, 1) * _SYMMETRY_NUM, dtype=np.float32)
        probability_labels = np.zeros((max(n, 1) * _SYMMETRY_NUM, _NUM_MOVES), dtype=np.float32)
        if n > 0:
            trainer.writeSamples(&sample_states[0, 0], &evaluation_labels[0], &probability_labels[0, 0])
        n *= _SYMMETRY_NUM
        return (np.asarray(sample_states)[:n], np.asarray(evaluation_labels)[:n],
                np.asarray(probability_labels)[:n], trainer.score(), trainer.avg_mate_length())
    finally:
        del trainer


def play_games_fused(int num_games, str log_folder, int seed, int max_searches, int searches_per_eval,
                     float c_puct, float epsilon, int num_logged, weights, int precision):
    """play_games with the evaluation on the device: `weights` is the folded 127 997-float vector
    (corintho_ai_b200.fold_batchnorm / tflite_import.load_tflite_weights), precision 0 fp32, 1 bf16,
    2 fp16, 3 bf16x3. Same return value as play_games."""
    cdef Trainer *trainer = new Trainer(num_games, log_folder.encode(), seed, max_searches, searches_per_eval,
                                        c_puct, epsilon, num_logged, 1, False)
    cdef float[::1] w = np.ascontiguousarray(weights, dtype=np.float32)
    cdef int n
    cdef float[:, ::1] sample_states
    cdef float[::1] evaluation_labels
    cdef float[:, ::1] probability_labels
    try:
        trainer.setWeights(0, &w[0], w.shape[0], precision)
        while not trainer.runSelfplay(0, 1):
            pass
        n = trainer.num_samples()
        sample_states = np.zeros((max(n, 1) * _SYMMETRY_NUM, _GAME_STATE_SIZE), dtype=np.float32)
        evaluation_labels = np.zeros(max(n, 1) * _SYMMETRY_NUM, dtype=np.float32)
        probability_labels = np.zeros((max(n, 1) * _SYMMETRY_NUM, _NUM_MOVES), dtype=np.float32)
        if n > 0:
            trainer.writeSamples(&sample_states[0, 0], &evaluation_labels[0], &probability_labels[0, 0])
        n *= _SYMMETRY_NUM
        return (np.asarray(sample_states)[:n], np.asarray(evaluation_labels)[:n],
                np.asarray(probability_labels)[:n], trainer.score(), trainer.avg_mate_length())
    finally:
        del trainer


cdef _get_tourney(Tourney *tourney, player_file, match_file):
    """tourney.pyx:63-110: the reference's player / match file formats."""
    model_ids = []
    with open(player_file, "r") as f:
        num_players = int(f.readline())
        for player_id in range(num_players):
            fields = f.readline().split()
            model_id, max_searches, searches_per_eval = int(fields[0]), int(fields[1]), int(fields[2])
            c_puct, epsilon, random = float(fields[3]), float(fields[4]), int(fields[5])
            tourney.addPlayer(player_id, model_id, max_searches, searches_per_eval, c_puct, epsilon,
                              random == 1)
            if model_id not in model_ids:
                model_ids.append(model_id)
    with open(match_file, "r") as f:
        num_matches = int(f.readline())
        for _ in range(num_matches):
            player1, player2, logging = [int(x) for x in f.readline().split()]
            tourney.addMatch(player1, player2, logging == 1)
    return model_ids


def run_tourney(evaluators, str player_file, str match_file, str log_folder, int num_threads, record=None):
    """tourney.pyx:176-205 (run) with play_games (112-173) inside: `evaluators[model_id](rows) ->
    (eval[n], probs[n,96])`. Writes log_folder/scores.txt; returns the number of passes over the
    model ids. `record`, if a list, receives (model_id, rows.copy()) per evaluation."""
    cdef Tourney *tourney = new Tourney(num_threads, log_folder.encode())
    cdef float[::1] ev
    cdef float[:, ::1] probs
    cdef float[:, ::1] game_states
    cdef int n, mid, rounds = 0
    try:
        model_ids = _get_tourney(tourney, player_file, match_file)
        ev = np.zeros(tourney.max_rows(), dtype=np.float32)
        probs = np.zeros((tourney.max_rows(), _NUM_MOVES), dtype=np.float32)
        game_states = np.zeros((tourney.max_rows(), _GAME_STATE_SIZE), dtype=np.float32)
        while not tourney.all_done():
            for mid in model_ids:
                n = tourney.num_requests(mid) if mid >= 0 else 0   # negative ids: random players
                if n > 0:
                    tourney.writeRequests(&game_states[0, 0], mid)
                    rows = np.asarray(game_states)[:n]
                    if record is not None:
                        record.append((mid, rows.copy()))
                    e, p = evaluators[mid](rows)
                    np.asarray(ev)[:n] = np.asarray(e, dtype=np.float32).reshape(-1)
                    np.asarray(probs)[:n] = p
                tourney.doIteration(&ev[0], &probs[0, 0], mid)
            rounds += 1
        tourney.writeScores(f"{log_folder}/scores.txt".encode())
        return rounds
    finally:
        del tourney


def trainer_rejects(int num_games, int searches_per_eval):
    """The error path: an invalid configuration surfaces as a Python exception via `except +`."""
    cdef Trainer *t = new Trainer(num_games, b"", 1, 16, searches_per_eval, 1.0, 0.25, 0, 1, False)
    del t
