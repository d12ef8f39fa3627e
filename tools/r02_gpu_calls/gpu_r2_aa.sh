#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_trainer.py tests/test_gpu_net.py tests/test_gpu_tourney.py -m gpu -x -q 2>&1 | tail -2
python tools/time_full.py 4096 800 3
CB200_GROUPS=1 CB200_NO_PERSISTENT=1 python tools/prof_selfplay.py 4096 800 300 bf16 2>&1 | grep -E "game_step"
export CB200_LIB=$PWD/corintho_ai_b200/libcorintho_b200_prof.so
CB200_GROUPS=1 CB200_NO_PERSISTENT=1 timeout 120 python tools/prof_timeline.py 1 800 bf16 2>&1 | sed -n 9,12p
