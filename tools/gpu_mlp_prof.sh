export CB200_LIB=$PWD/corintho_ai_b200/libcorintho_b200_prof.so
CB200_GROUPS=1 timeout 120 python tools/prof_selfplay.py 1 800 40 bf16 noprof 2>&1 | grep mlp_tc | tail -6
CB200_GROUPS=1 timeout 120 python tools/prof_selfplay.py 16 800 40 bf16 noprof 2>&1 | grep mlp_tc | tail -4
