#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/r2u_noinline.log; : > $out
timeout 300 python -m pytest tests/test_gpu_trainer.py tests/test_gpu_net.py -m gpu -x -q 2>&1 | tail -2 >> $out
for MB in 4 6 7 8; do
  echo "== minblocks $MB" >> $out
  CB200_MINBLOCKS=$MB CB200_GROUPS=1 CB200_NO_PERSISTENT=1 CB200_YIELD=0 python tools/prof_selfplay.py 4096 800 300 bf16 2>&1 | grep -E "game_step" >> $out
  CB200_MINBLOCKS=$MB python tools/time_full.py 4096 800 3 >> $out 2>&1
done
cat $out
