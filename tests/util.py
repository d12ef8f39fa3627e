"""Shared helpers for the test-suite (numpy only)."""
import hashlib

import numpy as np

M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def step_rnd(seed, n):
    """numpy twin of cb200::step_rnd (corintho_ai_b200/csrc/rules.cuh): per-state random word."""
    i = np.arange(n, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = np.uint64(seed) + (i + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(32)).astype(np.uint32)


def sha(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def run_trainer(trainer, evaluator, testing=False, allow_empty=False):
    """Drive any Trainer-like object through main.pyx:142-170's loop and digest everything the
    reference API exposes: per-round request counts + request rows, samples, score, mate length."""
    from oracle.pyoracle import play_out
    rec = []
    rounds = play_out(trainer, evaluator=evaluator, to_play=0 if testing else -1, record=rec,
                      allow_empty=allow_empty)
    counts = np.array([r.shape[0] for _, r in rec], np.int32)
    tps = np.array([tp for tp, _ in rec], np.int32)
    req_hash = sha(*[r for _, r in rec]) if rec else sha(np.zeros(0))
    out = {"rounds": rounds, "counts": counts, "to_play": tps, "req_hash": req_hash,
           "num_samples": trainer.num_samples(), "score": np.float32(trainer.score()),
           "mate": np.float32(trainer.avg_mate_length())}
    if not testing:
        gs, ev, pr = trainer.write_samples()
        out["samples"] = (gs, ev, pr)
        out["samples_hash"] = sha(gs, ev, pr)
    return out


TRAINER_GRID = [
    # (num_games, seed, max_searches, spe, c_puct, epsilon, testing)
    (1, 12345, 1, 1, 1.0, 0.25, False),
    (3, 12345, 2, 1, 1.0, 0.25, False),
    (3, 12345, 2, 2, 1.0, 0.25, False),
    (3, 12345, 16, 1, 1.0, 0.25, False),
    (3, 12345, 16, 16, 1.0, 0.25, False),
    (3, 12345, 96, 16, 1.0, 0.25, False),
    (3, 12345, 96, 96, 1.0, 0.25, False),
    (3, 12345, 400, 16, 1.0, 0.25, False),
    (1, 12345, 400, 400, 1.0, 0.25, False),
    (16, 7, 200, 16, 3.0, 0.25, False),
    (8, 99, 800, 16, 1.0, 0.0, False),
    (8, 5, 100, 8, 1.0, 0.25, True),
    (6, 21, 64, 16, 1.0, 0.0, True),
    (4, 11, 1600, 16, 3.0, 0.25, False),
]


def grid_key(cfg):
    g, seed, ms, spe, cp, eps, testing = cfg
    return f"g{g}_s{seed}_m{ms}_e{spe}_c{cp}_x{eps}_t{int(testing)}"
