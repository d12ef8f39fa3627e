mkdir -p gpurun_out
CB200_GROUPS=1 python tools/prof_selfplay.py 4096 800 300 bf16 noprof > gpurun_out/plain4.log 2>&1 &&
CB200_GROUPS=1 ncu --set full --clock-control none --import-source on -k regex:k_iterate -s 250 -c 1 -o gpurun_out/prof_iterate4 -f python tools/prof_selfplay.py 4096 800 300 bf16 noprof > gpurun_out/ncu4.log 2>&1
echo "ncu rc=$?"
