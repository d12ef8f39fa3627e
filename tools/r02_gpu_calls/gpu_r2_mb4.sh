#!/bin/bash
# new defaults (6 CTAs per SM for the lock-step game step; 8 stream groups from 16 384 games): tests, timing, bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python tools/time_full.py 4096 800 6 bf16
timeout 600 python tools/time_full.py 32768 800 2 bf16
timeout 600 python tools/time_full.py 8192 800 3 bf16
CB200_MINBLOCKS=4 timeout 600 python tools/time_full.py 8192 800 3 bf16
timeout 600 python tools/match_bench.py 2>&1 | tail -3
