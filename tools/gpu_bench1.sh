set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_net.py -q -m gpu --timeout 300 > gpurun_out/t_net.log 2>&1; echo "net rc=$?"; tail -5 gpurun_out/t_net.log
timeout 900 python bench.py --steps 2 --warmup 1 --ref-games 16 > gpurun_out/bench_fp32.json 2> gpurun_out/bench_fp32.err; echo "bench rc=$?"
cat gpurun_out/bench_fp32.json; tail -5 gpurun_out/bench_fp32.err
