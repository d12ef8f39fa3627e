"""Per-32-iteration timeline of a full fused self-play run: kernel times and live games."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import corintho_ai_b200 as cb
games = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
sims = int(sys.argv[2]) if len(sys.argv) > 2 else 800
prec = sys.argv[3] if len(sys.argv) > 3 else "bf16"
t = cb.Trainer(games, "", 12345, sims, 16, 1.0, 0.25)
t.set_weights(cb.fold_batchnorm(cb.random_weights(0)), 0, prec)
t.set_profiling(True)
t.phase_profile(True)
prev = {k: v["ms"] for k, v in t.kernel_times().items()}
it = 0; psims = 0
t0 = time.time()
while True:
    done = t.run_selfplay(64, stagger=False)
    it += 64
    kt = t.kernel_times(); c = t.counters()
    live = int((t.game_results() == 0).sum())
    d = {k: kt[k]["ms"] - prev[k] for k in kt}; prev = {k: kt[k]["ms"] for k in kt}
    print("iter %5d live %5d sims %9d  per-iter us: game_step %7.1f net %6.1f scan %5.1f pack %5.1f" % (
        it, live, c["simulations"] - psims, 1e3 * d["game_step"] / 64, 1e3 * d["network"] / 64, 1e3 * d["scan"] / 64, 1e3 * d["pack"] / 64))
    psims0 = psims
    psims = c["simulations"]
    pp = t.phase_profile(True).astype(np.float64)
    print("      max-warp kcycles (sum over 64 launches of per-launch max is not available; max over window): ingest %.0f search %.0f move %.0f | mean/warp-iter kcycles: ingest %.1f search %.1f move %.1f | rollbacks %d copied words %d" % (pp[0]/1e3, pp[1]/1e3, pp[2]/1e3, pp[3]/1e3/max(1,live*64), pp[4]/1e3/max(1,live*64), pp[5]/1e3/max(1,live*64), pp[6], pp[7]))
    if pp[10] > 0:
        print("      select %.0f cycles/level (%.2f levels/sim, exact pass %.2f%%), expand %.0f cycles/expansion (%d expansions), rollback searches %d" % (pp[8]/pp[10], pp[10]/max(1, c["simulations"]-psims0), 100*pp[12]/pp[10], pp[9]/max(1,pp[11]), pp[11], pp[6]))
    if done: break
print("total %.3f s" % (time.time() - t0), t.counters())
