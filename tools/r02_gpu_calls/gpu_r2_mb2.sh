#!/bin/bash
# register budget at configs[3] size (32768 games on one GPU) and on the match path
for mb in 4 6 7; do
  CB200_MINBLOCKS=$mb timeout 600 python tools/time_full.py 32768 800 2 bf16
done
for mb in 4 6; do
  CB200_MINBLOCKS=$mb timeout 600 python tools/time_full.py 1024 800 4 bf16
done
