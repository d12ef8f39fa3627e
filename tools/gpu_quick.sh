timeout 600 python -m pytest tests/test_gpu_net.py tests/test_gpu_trainer.py -q -m gpu --timeout 200 -x 2>&1 | tail -2
echo "== whole run (variants by live games)"; CB200_GROUPS=1 timeout 120 python tools/prof_selfplay.py 4096 800 0 bf16 noprof 2>&1 | grep done
echo "== whole run (fixed 64-reg variant)"; CB200_FIXED_VARIANT=1 CB200_GROUPS=1 timeout 120 python tools/prof_selfplay.py 4096 800 0 bf16 noprof 2>&1 | grep done
echo "== 1 game"; CB200_GROUPS=1 timeout 120 python tools/prof_selfplay.py 1 800 300 bf16 2>&1 | tail -1
echo "== 32768 games"; CB200_GROUPS=1 timeout 300 python tools/prof_selfplay.py 32768 800 0 bf16 noprof 2>&1 | grep done
