mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_net.py -m gpu -x -q -k "two_models" 2>&1 | tail -3
echo "== config[3] shape on one GPU: 32768 games x 800 sims"; timeout 600 python tools/prof_selfplay.py 32768 800 0 bf16 noprof 2>&1 | grep -E "done|Error|error" | tee gpurun_out/config3_1gpu.log
for np in 0 1; do
echo "== config[4] shape: 1250 games/GPU x 1600 sims, eps 0, testing (two nets), CB200_NO_PERSISTENT=$np"; CB200_NP=$np timeout 600 python - <<'PY' 2>&1 | tee gpurun_out/config4_1gpu_np$np.log
import sys, os, time
if os.environ["CB200_NP"] == "1": os.environ["CB200_NO_PERSISTENT"] = "1"
sys.path.insert(0, os.getcwd())
import corintho_ai_b200 as cb
t = cb.Trainer(1250, "", 7, 1600, 16, 1.0, 0.0, 0, 1, True)
t.set_weights(cb.fold_batchnorm(cb.random_weights(1)), 0, "bf16")
t.set_weights(cb.fold_batchnorm(cb.random_weights(2)), 1, "bf16")
t.run_selfplay(0)
best = 1e9
for r in range(3):
    t.reset(7)
    t0 = time.time(); done = t.run_selfplay(0); best = min(best, time.time() - t0)
c = t.counters()
print("done", done, "seconds %.3f" % best, c, "sims/s %.3e games/s %.1f score(model0) %.4f" % (c["simulations"]/best, 1250/best, float(t.score())))
PY
done
