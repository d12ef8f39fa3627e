mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_net.py tests/test_gpu_trainer.py -q -m gpu --timeout 200 -x > gpurun_out/t_all.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/t_all.log
for g in 1 8 16; do echo "== groups $g (no per-launch events)"; CB200_GROUPS=$g timeout 120 python tools/prof_selfplay.py 4096 800 0 bf16 noprof 2>&1 | grep done; done
echo "== dense phase, 1 group, 300 iterations"; CB200_GROUPS=1 timeout 120 python tools/prof_selfplay.py 4096 800 300 bf16 2>&1 | tail -3
CB200_GROUPS=1 timeout 120 python tools/prof_timeline.py 4096 800 bf16 2>&1 | head -4
