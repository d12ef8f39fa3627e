#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/r2s_sweeps.log; : > $out
for c in 2368 2072 1776 1480 1184; do CB200_PS_CAPACITY=$c python tools/time_full.py 4096 800 3 >> $out 2>&1; done
for y in 48 64 128 160; do CB200_YIELD=$y python tools/time_full.py 4096 800 3 >> $out 2>&1; done
for r in 25 35 65 75; do CB200_PS_REDEAL_PCT=$r python tools/time_full.py 4096 800 3 >> $out 2>&1; done
CB200_PS_REDEAL_MIN=16 python tools/time_full.py 4096 800 3 >> $out 2>&1
CB200_PS_REDEAL_MIN=256 python tools/time_full.py 4096 800 3 >> $out 2>&1
CB200_MINBLOCKS=6 CB200_GROUPS=8 python tools/time_full.py 4096 800 3 >> $out 2>&1
CB200_MINBLOCKS=6 CB200_GROUPS=12 python tools/time_full.py 4096 800 3 >> $out 2>&1
CB200_GROUPS=12 python tools/time_full.py 4096 800 3 >> $out 2>&1
cat $out
