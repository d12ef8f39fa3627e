#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_game_step.py -m gpu -x -q 2>&1 | tail -1
python tools/time_k1.py; python tools/time_k1.py
