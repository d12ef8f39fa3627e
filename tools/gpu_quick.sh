timeout 600 python -m pytest tests/test_gpu_net.py tests/test_gpu_trainer.py -q -m gpu --timeout 200 -x 2>&1 | tail -1
for i in 1 2; do echo "== whole run"; CB200_GROUPS=1 timeout 120 python tools/prof_selfplay.py 4096 800 0 bf16 noprof 2>&1 | grep done; done
echo "== dense"; CB200_GROUPS=1 timeout 120 python tools/prof_selfplay.py 4096 800 300 bf16 2>&1 | tail -1
echo "== 1 game"; CB200_GROUPS=1 timeout 120 python tools/prof_selfplay.py 1 800 300 bf16 2>&1 | tail -1
