set -x
mkdir -p gpurun_out
python tools/prof_selfplay.py 4096 800 300 fp32 > gpurun_out/prof_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_iterate -s 250 -c 2 -o gpurun_out/prof_iterate -f python tools/prof_selfplay.py 4096 800 300 fp32 > gpurun_out/ncu_iterate.log 2>&1
echo "ncu rc=$?"
cat gpurun_out/prof_plain.log
tail -5 gpurun_out/ncu_iterate.log
ls -la gpurun_out
