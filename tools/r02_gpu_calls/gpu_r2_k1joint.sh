#!/bin/bash
# the last GPU minute of the round: the branch-free joint finish (CB200_K1_PAIR=5 / 6) against the default on 64 Mi positions
mkdir -p gpurun_out
timeout 45 python tools/k1_pair_check.py > gpurun_out/r2k1j_variants.log 2>&1; echo "variants rc=$?"; tail -1 gpurun_out/r2k1j_variants.log
