mkdir -p gpurun_out
CB200_GROUPS=1 CB200_NO_PERSISTENT=1 python tools/prof_selfplay.py 4096 800 300 bf16 noprof > gpurun_out/plain_prof.log 2>&1 &&
CB200_GROUPS=1 CB200_NO_PERSISTENT=1 ncu --set full --clock-control none --import-source on -k regex:"k_iterate|k_mlp_tc" -s 500 -c 2 -o gpurun_out/prof_final -f python tools/prof_selfplay.py 4096 800 300 bf16 noprof > gpurun_out/ncu_final.log 2>&1
echo "full rc=$?"
ncu -i gpurun_out/prof_final.ncu-rep --page source --print-source cuda,sass --csv > /tmp/dense.csv 2>/dev/null
python tools/ncu_lines.py /tmp/dense.csv > gpurun_out/dense_lines.txt
ls -la gpurun_out | tail -4
