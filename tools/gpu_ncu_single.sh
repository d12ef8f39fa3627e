set -e
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'k_iterate|k_mlp_tc' -s 400 -c 16 -o /tmp/single_game -f \
  env CB200_GROUPS=1 python tools/prof_selfplay.py 1 800 300 bf16 noprof > gpurun_out/ncu_single.log 2>&1
ncu -i /tmp/single_game.ncu-rep --page source --print-source cuda,sass --csv > /tmp/single.csv 2>/dev/null
python tools/ncu_lines.py /tmp/single.csv > gpurun_out/single_lines.txt
ls -la gpurun_out/ /tmp/single.csv
