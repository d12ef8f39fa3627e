#!/bin/bash
# round 2, call D: lane-group / register-budget variants of the game step
mkdir -p gpurun_out
out=gpurun_out/r2d_variants.log
: > $out
for v in "32 4" "32 5" "32 6" "16 4" "16 3"; do
  set -- $v
  echo "== lanes $1 minblocks $2" >> $out
  CB200_LANES=$1 CB200_MINBLOCKS=$2 CB200_GROUPS=1 CB200_NO_PERSISTENT=1 python tools/prof_selfplay.py 4096 800 300 bf16 2>&1 | grep -E "game_step|network" >> $out
  CB200_LANES=$1 CB200_MINBLOCKS=$2 python tools/time_full.py 4096 800 3 >> $out 2>&1
  CB200_LANES=$1 CB200_MINBLOCKS=$2 CB200_NO_LIVE_LIST=1 python tools/time_full.py 4096 800 3 >> $out 2>&1
done
echo "== config3 shape" >> $out
for v in "32 4" "16 4" "16 3"; do
  set -- $v
  CB200_LANES=$1 CB200_MINBLOCKS=$2 python tools/time_full.py 32768 800 1 >> $out 2>&1
done
cat $out
