mkdir -p gpurun_out
python tools/prof_game_step.py > gpurun_out/plain_k1.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_game_step -s 12 -c 1 -o /tmp/prof_k1 -f python tools/prof_game_step.py > gpurun_out/ncu_k1.log 2>&1
echo "k1 rc=$?"
ncu -i /tmp/prof_k1.ncu-rep --page source --print-source cuda,sass --csv > /tmp/k1.csv 2>/dev/null
python tools/ncu_lines.py /tmp/k1.csv > gpurun_out/k1_lines.txt
