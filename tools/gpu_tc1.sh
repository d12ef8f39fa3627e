set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_net.py -q -m gpu --timeout 200 -s > gpurun_out/t_net.log 2>&1; echo "net rc=$?"; tail -40 gpurun_out/t_net.log
timeout 300 python tools/prof_selfplay.py 4096 800 300 bf16 > gpurun_out/prof_bf16.log 2>&1; cat gpurun_out/prof_bf16.log
