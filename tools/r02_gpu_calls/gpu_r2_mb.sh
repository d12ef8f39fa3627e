#!/bin/bash
# register budget of the lock-step game step once more on the final build: 4 / 5 / 6 CTAs per SM, 6 runs each
for mb in 4 5 6 4 5 6; do
  CB200_MINBLOCKS=$mb timeout 300 python tools/time_full.py 4096 800 6 bf16
done
