#!/bin/bash
mkdir -p gpurun_out
export CB200_LIB=$PWD/corintho_ai_b200/libcorintho_b200_prof.so
CB200_GROUPS=1 CB200_NO_PERSISTENT=1 timeout 120 python tools/prof_timeline.py 1 800 bf16 2>&1 | sed -n 9,16p
CB200_GROUPS=1 CB200_NO_PERSISTENT=1 timeout 120 python tools/prof_timeline.py 4096 800 bf16 2>&1 | sed -n 17,20p
