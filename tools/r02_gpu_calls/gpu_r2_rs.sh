#!/bin/bash
# root slot array in shared memory during a doIteration's searches: parity + timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_trainer.py tests/test_gpu_tourney.py tests/test_gpu_bench_config.py -x -q 2>&1 | tail -5
timeout 300 python tools/time_full.py 4096 800 3 bf16
CB200_GROUPS=1 CB200_NO_PERSISTENT=1 timeout 300 python tools/prof_selfplay.py 4096 800 300 bf16 2>&1 | grep -E "game_step|network" | head -4
timeout 300 python tools/prof_selfplay.py 592 800 300 bf16 2>&1 | grep -E "game_step|network|fused" | head -4
