mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
CB200_GROUPS=1 CB200_NO_PERSISTENT=1 python tools/prof_selfplay.py 4096 800 300 bf16 noprof > gpurun_out/plain_prof.log 2>&1 &&
CB200_GROUPS=1 CB200_NO_PERSISTENT=1 ncu --set full --clock-control none --import-source on -k regex:"k_iterate|k_mlp_tc" -s 500 -c 2 -o gpurun_out/prof_final -f python tools/prof_selfplay.py 4096 800 300 bf16 noprof > gpurun_out/ncu_final.log 2>&1
echo "full rc=$?"
