#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_game_step.py tests/test_gpu_edge.py -m gpu -x -q > gpurun_out/r2p_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2p_tests.log; tail -5 gpurun_out/r2p_tests.log
python tools/time_k1.py > gpurun_out/r2p_k1_split.log 2>&1; cat gpurun_out/r2p_k1_split.log
CB200_K1_SELECT_ONLY=1 python tools/time_k1.py > gpurun_out/r2p_k1_select.log 2>&1; cat gpurun_out/r2p_k1_select.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_game_step" -s 12 -c 1 -o gpurun_out/r2p_k1 -f python tools/time_k1.py > gpurun_out/r2p_ncu_k1.log 2>&1
ncu -i gpurun_out/r2p_k1.ncu-rep --page raw --csv > gpurun_out/r2p_k1_raw.csv 2>/dev/null
ncu -i gpurun_out/r2p_k1.ncu-rep --page source --print-source cuda,sass --csv > /tmp/k1.csv 2>/dev/null
python tools/ncu_lines.py /tmp/k1.csv > gpurun_out/r2p_k1_lines.txt
rm -f gpurun_out/r2p_k1.ncu-rep
