"""Small end-to-end run of every kernel and execution mode (fused lock-step, parking, persistent
tail, two-model persistent, external evaluator, fp32 network, tourney, text logs): a quick
sanity run on a GPU box, also suitable for compute-sanitizer where that is available."""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import corintho_ai_b200 as cb
from oracle.pyoracle import play_out, play_tourney

flat = cb.fold_batchnorm(cb.random_weights(1))
flat2 = cb.fold_batchnorm(cb.random_weights(2))
d = tempfile.mkdtemp()
# fused training run: lock-step, parking, hand-over to the persistent kernel, logs
os.environ["CB200_YIELD"] = "12"; os.environ["CB200_YIELD_MIN_LIVE"] = "1"
t = cb.Trainer(20, d, 3, 24, 8, 1.0, 0.25, 2)
t.set_weights(flat, 0, "bf16")
assert t.run_selfplay(0, stagger=True)
print("fused", t.counters(), t.num_samples())
t.write_samples()
# two-model persistent
u = cb.Trainer(12, "", 4, 16, 4, 1.0, 0.0, 0, 1, True)
u.set_weights(flat, 0, "fp16"); u.set_weights(flat2, 1, "fp16")
assert u.run_selfplay(0)
print("two-model", u.counters(), float(u.score()))
# external-evaluator mode + fp32 network + K1
v = cb.Trainer(6, "", 5, 16, 4, 1.0, 0.25)
play_out(v)
print("external", v.counters())
w = cb.Trainer(6, "", 5, 16, 4, 1.0, 0.25)
w.set_weights(flat, 0, "fp32")
assert w.run_selfplay(0, stagger=False)
# tourney
T = cb.Tourney(1, "")
T.addPlayer(0, 0, 16, 4, 1.0, 0.25); T.addPlayer(1, 1, 12, 3, 2.0, 0.0); T.addPlayer(2, -1, 1, 1, 1.0, 0.25, True)
for a, b in [(0, 1), (1, 0), (2, 0), (1, 2), (2, 2)]:
    T.addMatch(a, b)
play_tourney(T)
print("tourney", T.scores())
print("OK")
