set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_net.py tests/test_gpu_trainer.py -q -m gpu --timeout 200 -x > gpurun_out/t_all.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/t_all.log
for g in 1 2 4 8 16 32; do echo "== groups $g"; CB200_GROUPS=$g timeout 120 python tools/prof_selfplay.py 4096 800 0 bf16 2>&1 | tail -6; done
