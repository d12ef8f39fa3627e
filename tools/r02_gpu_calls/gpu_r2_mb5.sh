#!/bin/bash
# persistent take-over threshold and yield budget once more with the 6-CTA game step
for cap in 1184 1776 2072; do
  CB200_PS_CAPACITY=$cap timeout 300 python tools/time_full.py 4096 800 6 bf16
done
for y in 64 96 128; do
  CB200_YIELD=$y timeout 300 python tools/time_full.py 4096 800 6 bf16
done
CB200_GROUPS=4 timeout 300 python tools/time_full.py 4096 800 6 bf16
CB200_GROUPS=8 timeout 300 python tools/time_full.py 4096 800 6 bf16
