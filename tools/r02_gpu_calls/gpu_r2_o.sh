#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/r2o_variants.log; : > $out
for MB in 4 6 7 8; do
  echo "== minblocks $MB" >> $out
  CB200_MINBLOCKS=$MB CB200_GROUPS=1 CB200_NO_PERSISTENT=1 CB200_YIELD=0 python tools/prof_selfplay.py 4096 800 300 bf16 2>&1 | grep -E "game_step" >> $out
  CB200_MINBLOCKS=$MB python tools/time_full.py 4096 800 3 >> $out 2>&1
  CB200_MINBLOCKS=$MB CB200_GROUPS=1 python tools/time_full.py 4096 800 3 >> $out 2>&1
  CB200_MINBLOCKS=$MB CB200_GROUPS=2 python tools/time_full.py 4096 800 3 >> $out 2>&1
done
cat $out
