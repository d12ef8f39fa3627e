"""Where the end-to-end step of bench.py spends its time (host-visible phases)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import corintho_ai_b200 as cb
G = 4096
flat = cb.fold_batchnorm(cb.random_weights(0))
pinned_w = torch.from_numpy(flat).pin_memory()
tr = cb.Trainer(G, "", 12345, 800, 16, 1.0, 0.25)
tr.set_weights(flat, 0, "bf16")
tr.run_selfplay(0, stagger=False)
cap = int(tr.num_samples() * 8 * 1.25) + 1024
gs = torch.empty((cap, 70), dtype=torch.float32).pin_memory().numpy()
ev = torch.empty((cap,), dtype=torch.float32).pin_memory().numpy()
pr = torch.empty((cap, 96), dtype=torch.float32).pin_memory().numpy()
n = tr.num_samples(); tr.writeSamples(gs[:n*8], ev[:n*8], pr[:n*8])
for rep in range(3):
    t = [time.perf_counter()]
    tr.set_weights(pinned_w.numpy(), 0, "bf16"); torch.cuda.synchronize(); t.append(time.perf_counter())
    tr.reset(3000 + rep); torch.cuda.synchronize(); t.append(time.perf_counter())
    tr.run_selfplay(0, stagger=False); torch.cuda.synchronize(); t.append(time.perf_counter())
    n = tr.num_samples(); t.append(time.perf_counter())
    tr.writeSamples(gs[:n*8], ev[:n*8], pr[:n*8]); torch.cuda.synchronize(); t.append(time.perf_counter())
    d = [1e3 * (b - a) for a, b in zip(t, t[1:])]
    print("set_weights %.2f ms | reset %.2f | run %.2f | num_samples %.2f | writeSamples %.2f (%.0f MB) | total %.2f" % (
        d[0], d[1], d[2], d[3], d[4], (gs[:n*8].nbytes + ev[:n*8].nbytes + pr[:n*8].nbytes) / 1e6, sum(d)))
