export CB200_LIB=$PWD/corintho_ai_b200/libcorintho_b200_prof.so
CB200_GROUPS=1 timeout 120 python tools/prof_timeline.py 4096 800 bf16 > gpurun_out/timeline.txt 2>&1
