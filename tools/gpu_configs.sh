mkdir -p gpurun_out
echo "== config[3] shape on one GPU: 32768 games x 800 sims"; CB200_GROUPS=1 timeout 600 python tools/prof_selfplay.py 32768 800 0 bf16 noprof 2>&1 | grep -E "done|Error|error" | tee gpurun_out/config3_1gpu.log
echo "== config[4] shape: 1250 games/GPU x 1600 sims, eps 0, testing (two nets)"; timeout 600 python - <<'PY' 2>&1 | tee gpurun_out/config4_1gpu.log
import sys, os, time
sys.path.insert(0, os.getcwd())
import corintho_ai_b200 as cb
t = cb.Trainer(1250, "", 7, 1600, 16, 1.0, 0.0, 0, 1, True)
t.set_weights(cb.fold_batchnorm(cb.random_weights(1)), 0, "bf16")
t.set_weights(cb.fold_batchnorm(cb.random_weights(2)), 1, "bf16")
t.run_selfplay(3)
t.reset(7)
t0 = time.time(); done = t.run_selfplay(0); dt = time.time() - t0
c = t.counters()
print("done", done, "seconds %.3f" % dt, c, "sims/s %.3e games/s %.1f score(model0) %.4f" % (c["simulations"]/dt, 1250/dt, float(t.score())))
PY
echo "== e2e check"; timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_e2e.json 2>gpurun_out/bench_e2e.err; python -c "
import json; d=json.load(open('gpurun_out/bench_e2e.json')); print(d['value'], d['e2e'])"
