#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_game_step.py tests/test_gpu_edge.py -m gpu -x -q 2>&1 | tail -2
python tools/time_k1.py; python tools/time_k1.py
