#!/bin/bash
# single-tile network CTA capped at 104 / 96 registers so that it fits beside five 80-register game-step CTAs
for v in default n104 n96 default n104 n96; do
  if [ $v = default ]; then unset CB200_LIB; else export CB200_LIB=$PWD/corintho_ai_b200/libcorintho_b200_$v.so; fi
  echo "== $v"
  timeout 300 python tools/time_full.py 4096 800 6 bf16
done
for v in default n104; do
  if [ $v = default ]; then unset CB200_LIB; else export CB200_LIB=$PWD/corintho_ai_b200/libcorintho_b200_$v.so; fi
  timeout 600 python tools/time_full.py 32768 800 2 bf16
done
