#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tourney.py tests/test_gpu_logs.py -m gpu -x -q 2>&1 | tail -2
timeout 600 python tools/tourney_bench.py 8 16 400 cmp
timeout 600 python tools/tourney_bench.py 16 8 800
