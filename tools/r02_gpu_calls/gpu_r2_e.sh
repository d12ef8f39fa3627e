#!/bin/bash
# round 2, call E: defaults back to one warp per game; tail variants; full tests
mkdir -p gpurun_out
out=gpurun_out/r2e_variants.log
: > $out
python tools/time_full.py 4096 800 3 >> $out 2>&1
CB200_PS_LANES=16 python tools/time_full.py 4096 800 3 >> $out 2>&1
CB200_MINBLOCKS=5 python tools/time_full.py 4096 800 3 >> $out 2>&1
CB200_NO_LIVE_LIST=1 python tools/time_full.py 4096 800 3 >> $out 2>&1
CB200_ARENA_BUDGET_MB=90000 python tools/time_full.py 4096 800 3 >> $out 2>&1
CB200_GROUPS=4 python tools/time_full.py 4096 800 3 >> $out 2>&1
CB200_GROUPS=8 python tools/time_full.py 4096 800 3 >> $out 2>&1
cat $out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2e_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2e_tests.log
tail -5 gpurun_out/r2e_tests.log
CB200_LANES=16 CB200_PS_LANES=16 timeout 900 python -m pytest tests/test_gpu_trainer.py tests/test_gpu_net.py tests/test_gpu_bench_config.py -m gpu -x -q > gpurun_out/r2e_tests16.log 2>&1
echo "tests16 rc=$?" >> gpurun_out/r2e_tests16.log
tail -5 gpurun_out/r2e_tests16.log
