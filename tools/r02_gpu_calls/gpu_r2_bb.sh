#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/r2bb_misc.log; : > $out
echo "== network time per launch, 4096 games (65536 positions), single group lock-step" >> $out
for p in bf16 fp16 bf16x3 fp32; do echo "-- $p" >> $out; CB200_GROUPS=1 CB200_NO_PERSISTENT=1 python tools/prof_selfplay.py 4096 800 100 $p 2>&1 | grep -E "network" >> $out; done
echo "== full run with bf16x3 (lock-step only)" >> $out
python tools/time_full.py 4096 800 2 bf16x3 >> $out 2>&1
echo "== fused tourney" >> $out
timeout 600 python tools/tourney_bench.py 8 16 400 cmp >> $out 2>&1
timeout 600 python tools/tourney_bench.py 16 8 800 >> $out 2>&1
cat $out
