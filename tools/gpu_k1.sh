timeout 600 python -m pytest tests/test_gpu_game_step.py tests/test_gpu_trainer.py -q -m gpu --timeout 200 -x 2>&1 | tail -2
python - <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import torch, corintho_ai_b200 as cb
import bench
print(bench.measure_game_logic(torch, cb, torch.device("cuda", 0)))
PY
for g in 1 2 4; do echo "== groups $g"; CB200_GROUPS=$g timeout 120 python tools/prof_selfplay.py 4096 800 0 bf16 noprof 2>&1 | grep done; done
