#!/bin/bash
# closing check of the default form (CB200_K1_PAIR=6): golden vectors and every tail shape against the oracle
mkdir -p gpurun_out
timeout 14 python -m pytest tests/test_gpu_game_step.py -x -q -k "golden or ragged" > gpurun_out/r2k1f_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r2k1f_tests.log
