#!/bin/bash
# path prefetch in the backup phase of receive_eval: off / L1 / L2
mkdir -p gpurun_out
for v in pf0 pf1 default; do
  if [ $v = default ]; then unset CB200_LIB; else export CB200_LIB=$PWD/corintho_ai_b200/libcorintho_b200_$v.so; fi
  echo "== $v"
  timeout 300 python tools/time_full.py 4096 800 3 bf16
  CB200_GROUPS=1 CB200_NO_PERSISTENT=1 timeout 300 python tools/prof_selfplay.py 4096 800 300 bf16 2>&1 | grep -E "game_step" | head -1
done
unset CB200_LIB
timeout 900 python -m pytest tests/test_gpu_trainer.py tests/test_gpu_tourney.py tests/test_gpu_bench_config.py -x -q 2>&1 | tail -3
