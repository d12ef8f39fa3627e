set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 300 python -m pytest tests/test_gpu_game_step.py -q -m gpu --timeout 250 > gpurun_out/t_step.log 2>&1; echo "step rc=$?"
timeout 900 python -m pytest tests/test_gpu_trainer.py -q -m gpu --timeout 300 -x > gpurun_out/t_trainer.log 2>&1; echo "trainer rc=$?"
timeout 600 python -m pytest tests/test_gpu_net.py -q -m gpu --timeout 300 > gpurun_out/t_net.log 2>&1; echo "net rc=$?"
tail -5 gpurun_out/t_step.log; tail -30 gpurun_out/t_trainer.log; tail -30 gpurun_out/t_net.log
