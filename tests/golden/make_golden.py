#!/usr/bin/env python
"""Generate the committed golden fixtures from the COMPILED REFERENCE (oracle/_ref, i.e. the
unmodified sources under /root/reference). Run in the build container only:

    make -C oracle ref && python tests/golden/make_golden.py

Outputs (small, committed):
  tests/golden/rules.npz    20k reachable states (reference bit order) with the reference's
                            legal masks, flags, next states and NN encodings
  tests/golden/trainer.npz  per-config digests of full Trainer runs under the synthetic
                            evaluator oracle.pyoracle.synth_eval: request counts per round,
                            sha256 of every request row, samples (hash; full arrays for the
                            small configs), score, avg mate length
  tests/golden/tourney.npz  digests of full Tourney runs (per-player budgets, random players):
                            model id and row count of every evaluation, sha256 of all request
                            rows, final scores
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from oracle.pyoracle import RefLib, synth_eval  # noqa: E402
from util import (TOURNEY_CASES, TRAINER_GRID, grid_key, make_tourney, run_tourney, run_trainer,  # noqa: E402
                  step_rnd)


def main():
    R = RefLib()
    n, seed = 20000, 2024
    states = R.gen_states(777, n)
    rnd = step_rnd(seed, n)
    masks, flags, nxt, enc = R.step_batch(states, rnd)
    np.savez_compressed(os.path.join(HERE, "rules.npz"), states=states, seed=np.uint64(seed),
                        masks=masks, flags=flags, next=nxt, enc=enc.astype(np.float16))
    out = {}
    for cfg in TRAINER_GRID:
        g, s, ms, spe, cp, eps, testing = cfg
        t = R.trainer(num_games=g, seed=s, max_searches=ms, searches_per_eval=spe, c_puct=cp,
                      epsilon=eps, testing=testing)
        r = run_trainer(t, synth_eval, testing)
        k = grid_key(cfg)
        out[k + "/rounds"] = np.int64(r["rounds"])
        out[k + "/counts"] = r["counts"]
        out[k + "/to_play"] = r["to_play"]
        out[k + "/req_hash"] = np.bytes_(r["req_hash"])
        out[k + "/num_samples"] = np.int64(r["num_samples"])
        out[k + "/score"] = r["score"]
        out[k + "/mate"] = r["mate"]
        if not testing:
            out[k + "/samples_hash"] = np.bytes_(r["samples_hash"])
            if r["num_samples"] <= 64:
                gs, ev, pr = r["samples"]
                out[k + "/gs"], out[k + "/ev"], out[k + "/pr"] = gs, ev, pr
        print(k, r["rounds"], r["num_samples"], r["score"])
    np.savez_compressed(os.path.join(HERE, "trainer.npz"), **out)
    # Tourney / Match transcripts (tourney.h:12-46) under the same synthetic evaluator
    tout = {}
    for name in TOURNEY_CASES:
        r = run_tourney(make_tourney(R, name))
        tout[name + "/rounds"] = np.int64(r["rounds"])
        tout[name + "/models"] = r["models"]
        tout[name + "/counts"] = r["counts"]
        tout[name + "/req_hash"] = np.frombuffer(r["req_hash"].encode(), np.uint8)
        tout[name + "/scores"] = r["scores"]
        print("tourney", name, r["rounds"], "rounds,", len(r["counts"]), "evaluations", r["scores"][:, 2])
    np.savez_compressed(os.path.join(HERE, "tourney.npz"), **tout)


if __name__ == "__main__":
    main()
