#!/bin/bash
# parking inside the persistent kernels?
python tools/time_full.py 4096 800 3
for y in 64 96 128 192; do for m in 1184 512 128; do CB200_PS_YIELD=$y CB200_PS_YIELD_MIN_LIVE=$m python tools/time_full.py 4096 800 3; done; done
CB200_PS_YIELD=96 CB200_PS_YIELD_MIN_LIVE=512 timeout 300 python -m pytest tests/test_gpu_bench_config.py -m gpu -x -q 2>&1 | tail -2
