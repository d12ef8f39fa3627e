"""Device-timed full fused self-play runs (the bench.py step) under the current environment:
   python tools/time_full.py [games] [sims] [runs] [precision]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import corintho_ai_b200 as cb

games = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
sims = int(sys.argv[2]) if len(sys.argv) > 2 else 800
runs = int(sys.argv[3]) if len(sys.argv) > 3 else 3
prec = sys.argv[4] if len(sys.argv) > 4 else "bf16"
t = cb.Trainer(games, "", 12345, sims, 16, 1.0, 0.25)
t.set_weights(cb.fold_batchnorm(cb.random_weights(0)), 0, prec)
t.reset(999)
t.run_selfplay(0, stagger=False)
ms = []
for k in range(runs):
    t.reset(2000 + k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t.run_selfplay(0, stagger=False)
    e1.record()
    torch.cuda.synchronize()
    ms.append(e0.elapsed_time(e1))
c = t.counters()
tag = " ".join("%s=%s" % (k, v) for k, v in sorted(os.environ.items()) if k.startswith("CB200_"))
print("full run [%s] games %d: ms %s  sims/s %.3e (last run: %d sims, %d iterations)"
      % (tag, games, " ".join("%.1f" % m for m in ms), c["simulations"] / (ms[-1] * 1e-3), c["simulations"], c["iterations"]))
