export CB200_LIB=$PWD/corintho_ai_b200/libcorintho_b200_prof.so
CB200_GROUPS=1 CB200_NO_PERSISTENT=1 timeout 120 python tools/prof_selfplay.py 4096 800 60 bf16 noprof 2>&1 | grep tc_forward | tail -8
