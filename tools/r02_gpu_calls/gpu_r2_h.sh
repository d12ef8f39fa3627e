#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -s > gpurun_out/r2h_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2h_tests.log
grep -E "fixture|trained net|passed|failed|rc=|Error|error" gpurun_out/r2h_tests.log | tail -20
python tools/time_full.py 4096 800 3 > gpurun_out/r2h_time.log 2>&1; cat gpurun_out/r2h_time.log
CB200_GROUPS=1 CB200_NO_PERSISTENT=1 python tools/prof_selfplay.py 4096 800 300 bf16 2>&1 | grep -E "network|game_step" 
