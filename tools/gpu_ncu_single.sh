set -e
CB200_GROUPS=1 timeout 120 python tools/prof_selfplay.py 1 800 300 bf16 noprof | head -3
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_iterate -s 250 -c 1 -o gpurun_out/single_game -f \
  env CB200_GROUPS=1 python tools/prof_selfplay.py 1 800 300 bf16 noprof > gpurun_out/ncu_single.log 2>&1
ls -la gpurun_out/
