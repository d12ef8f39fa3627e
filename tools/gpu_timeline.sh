CB200_GROUPS=1 timeout 120 python tools/prof_timeline.py 4096 800 bf16 2>&1 | grep "^iter" > gpurun_out/timeline.txt
