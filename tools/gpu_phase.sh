timeout 600 python -m pytest tests/test_gpu_trainer.py -m gpu -x -q 2>&1 | tail -2
export CB200_LIB=$PWD/corintho_ai_b200/libcorintho_b200_prof.so
CB200_GROUPS=1 timeout 120 python tools/prof_timeline.py 4096 800 bf16 2>&1 | awk 'NR<=3 || (NR>=43 && NR<=45) || NR>=68'
echo "=== single game"
CB200_GROUPS=1 timeout 120 python tools/prof_timeline.py 1 800 bf16 2>&1 | head -6
