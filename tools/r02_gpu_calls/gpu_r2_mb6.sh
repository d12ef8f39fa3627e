#!/bin/bash
# yield budget of the lock-step loop, finer, with the 6-CTA game step
for y in 56 64 72 80 64 96; do
  CB200_YIELD=$y timeout 300 python tools/time_full.py 4096 800 6 bf16
done
CB200_YIELD=64 timeout 600 python tools/time_full.py 32768 800 2 bf16
CB200_YIELD=64 timeout 600 python tools/time_full.py 8192 800 3 bf16
