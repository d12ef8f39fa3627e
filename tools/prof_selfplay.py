"""Bounded fused self-play run for ncu / timing experiments:
   python tools/prof_selfplay.py [games] [sims] [iterations] [precision]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import corintho_ai_b200 as cb

games = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
sims = int(sys.argv[2]) if len(sys.argv) > 2 else 800
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 300  # 0 = to completion
prec = sys.argv[4] if len(sys.argv) > 4 else "fp32"
prof = not (len(sys.argv) > 5 and sys.argv[5] == "noprof")
t = cb.Trainer(games, "", 12345, sims, 16, 1.0, 0.25)
t.set_weights(cb.fold_batchnorm(cb.random_weights(0)), 0, prec)
t.run_selfplay(3, stagger=False)  # warm-up: module load, first launches
t.reset(12345)
t.set_profiling(prof)
t0 = time.time()
done = t.run_selfplay(iters, stagger=False)
dt = time.time() - t0
c = t.counters()
print("done", done, "seconds %.3f" % dt, c, "sims/s %.3e" % (c["simulations"] / dt))
for k, v in t.kernel_times().items():
    print("  %-10s %8.2f ms  %6d launches  %8.1f us/launch" % (k, v["ms"], v["launches"], 1e3 * v["ms"] / max(1, v["launches"])))
