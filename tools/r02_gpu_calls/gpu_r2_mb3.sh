#!/bin/bash
# configs[3] size: 5 CTAs per SM, and the stream-group count with 6
CB200_MINBLOCKS=5 timeout 600 python tools/time_full.py 32768 800 2 bf16
for g in 3 4 8 12; do
  CB200_GROUPS=$g CB200_MINBLOCKS=6 timeout 600 python tools/time_full.py 32768 800 2 bf16
done
CB200_LANES=16 CB200_MINBLOCKS=3 timeout 600 python tools/time_full.py 32768 800 2 bf16
CB200_LANES=16 CB200_MINBLOCKS=4 timeout 600 python tools/time_full.py 32768 800 2 bf16
