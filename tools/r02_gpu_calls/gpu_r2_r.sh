#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_net.py -m gpu -x -q > gpurun_out/r2r_tests_net.log 2>&1
echo "net tests rc=$?" >> gpurun_out/r2r_tests_net.log; tail -6 gpurun_out/r2r_tests_net.log
grep -q "rc=0" gpurun_out/r2r_tests_net.log || exit 1
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2r_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2r_tests.log; tail -4 gpurun_out/r2r_tests.log
CB200_GROUPS=1 CB200_NO_PERSISTENT=1 timeout 120 python tools/prof_selfplay.py 4096 800 300 bf16 2>&1 | grep -E "network|game_step"
timeout 120 python tools/time_full.py 4096 800 3
CB200_NO_OVERLAP=1 timeout 120 python tools/time_full.py 4096 800 3
