#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_trainer.py tests/test_gpu_net.py -m gpu -x -q 2>&1 | tail -2
python tools/time_full.py 4096 800 3
CB200_GROUPS=1 CB200_NO_PERSISTENT=1 python tools/prof_selfplay.py 4096 800 300 bf16 2>&1 | grep -E "game_step"
CB200_GROUPS=1 CB200_NO_PERSISTENT=1 python tools/prof_selfplay.py 1 800 300 bf16 2>&1 | grep -E "game_step"
python tools/prof_selfplay.py 1 800 300 bf16 2>&1 | grep -E "fused|seconds"
