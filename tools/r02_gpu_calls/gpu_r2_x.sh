#!/bin/bash
mkdir -p gpurun_out
export CB200_LIB=$PWD/corintho_ai_b200/libcorintho_b200_prof.so
echo "=== single game, lock-step" > gpurun_out/r2x_phase.log
CB200_GROUPS=1 CB200_NO_PERSISTENT=1 timeout 120 python tools/prof_timeline.py 1 800 bf16 2>&1 | sed -n 1,18p >> gpurun_out/r2x_phase.log
echo "=== 592 games, lock-step" >> gpurun_out/r2x_phase.log
CB200_GROUPS=1 CB200_NO_PERSISTENT=1 timeout 120 python tools/prof_timeline.py 592 800 bf16 2>&1 | sed -n 1,12p >> gpurun_out/r2x_phase.log
echo "=== 4096 games, lock-step" >> gpurun_out/r2x_phase.log
CB200_GROUPS=1 CB200_NO_PERSISTENT=1 timeout 120 python tools/prof_timeline.py 4096 800 bf16 2>&1 | sed -n 10,18p >> gpurun_out/r2x_phase.log
cat gpurun_out/r2x_phase.log
