#!/usr/bin/env python
"""Fused tourney throughput: a round-robin field of P players (one random-init network each, the
rating path of rating/round.py -> tourney.pyx) with M matches per ordered pair, played on the
device by cb200_tourney_run, and the same field through the external-evaluator protocol for
comparison (host round trips per model and iteration).  python tools/tourney_bench.py [P] [M] [sims]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import corintho_ai_b200 as cb

P = int(sys.argv[1]) if len(sys.argv) > 1 else 8
M = int(sys.argv[2]) if len(sys.argv) > 2 else 16
sims = int(sys.argv[3]) if len(sys.argv) > 3 else 400
weights = [cb.fold_batchnorm(cb.random_weights(200 + p)) for p in range(P)]


def build():
    t = cb.Tourney(1, "")
    for p in range(P):
        t.addPlayer(p, p, sims, 16, 1.0, 0.25)
    n = 0
    for a in range(P):
        for b in range(P):
            if a != b:
                for _ in range(M):
                    t.addMatch(a, b)
                    n += 1
    return t, n


t, n = build()
for p in range(P):
    t.set_weights(p, weights[p], "bf16")
t0 = time.perf_counter()
assert t.run(0)
dt = time.perf_counter() - t0
c = t.counters()
print("fused tourney: %d players, %d matches, %d sims/move: %.3f s = %.0f matches/s, %.3e sims/s, %d rounds"
      % (P, n, sims, dt, n / dt, c["simulations"] / dt, c["iterations"]))
fused_scores = t.scores()
if len(sys.argv) > 4:  # the same field through the external protocol (engine networks as evaluators)
    u, _ = build()
    helpers = []
    for p in range(P):
        h = cb.Trainer(1024, "", 1, 16, 16)
        h.set_weights(weights[p], 0, "bf16")
        helpers.append(h)
    rows = max(u.max_rows, 1)
    ev = np.zeros(rows, np.float32); pr = np.zeros((rows, 96), np.float32)
    t0 = time.perf_counter()
    while not u.all_done():
        for mid in u.model_ids:
            k = u.num_requests(mid)
            if k:
                req = u.write_requests(mid)
                for r0 in range(0, k, 16384):
                    e, q = helpers[mid].evaluate(req[r0:r0 + 16384])
                    ev[r0:r0 + len(e)], pr[r0:r0 + len(e)] = e, q
            u.doIteration(ev, pr, mid)
    dt2 = time.perf_counter() - t0
    print("external-evaluator protocol, same field: %.3f s = %.0f matches/s (%.1fx slower); scores equal: %s"
          % (dt2, n / dt2, dt2 / dt, u.scores() == fused_scores))
