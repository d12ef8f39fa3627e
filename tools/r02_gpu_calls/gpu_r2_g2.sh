#!/bin/bash
# stream groups / register budget once more with the 77 KB network CTA
for g in 4 8 12; do CB200_GROUPS=$g timeout 300 python tools/time_full.py 4096 800 6 bf16; done
CB200_MINBLOCKS=5 timeout 300 python tools/time_full.py 4096 800 6 bf16
CB200_MINBLOCKS=4 timeout 300 python tools/time_full.py 4096 800 6 bf16
CB200_MINBLOCKS=7 timeout 300 python tools/time_full.py 4096 800 6 bf16
timeout 300 python tools/time_full.py 4096 800 6 bf16
