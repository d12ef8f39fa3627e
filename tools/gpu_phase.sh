CB200_GROUPS=1 timeout 120 python tools/prof_timeline.py 4096 800 bf16 2>&1 | awk 'NR<=9 || (NR>=40 && NR<=48) || NR>=62'
echo "=== single game"
CB200_GROUPS=1 timeout 120 python tools/prof_timeline.py 1 800 bf16 2>&1 | head -12
