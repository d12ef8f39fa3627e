#!/bin/bash
# the paired K1 kernel: parity tests, full-size comparison with the earlier kernels, timings, smoke, ncu
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_game_step.py -x -q > gpurun_out/r2k1_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r2k1_tests.log
timeout 150 python tools/k1_pair_check.py > gpurun_out/r2k1_variants.log 2>&1; echo "variants rc=$?"; tail -1 gpurun_out/r2k1_variants.log
timeout 150 ncu --set full --clock-control none --import-source on -k regex:k_game_step_pairw -s 12 -c 1 -o /tmp/k1pair -f python tools/prof_game_step.py > gpurun_out/r2k1_ncu.log 2>&1; echo "ncu rc=$?"
ncu -i /tmp/k1pair.ncu-rep --page raw --csv > gpurun_out/r2k1_ncu_raw.csv 2>/dev/null
ncu -i /tmp/k1pair.ncu-rep --page source --print-source cuda,sass --csv > /tmp/k1pair_src.csv 2>/dev/null && python tools/ncu_lines.py /tmp/k1pair_src.csv > gpurun_out/r2k1_ncu_lines.txt 2>/dev/null
echo done
