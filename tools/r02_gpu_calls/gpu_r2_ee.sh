#!/bin/bash
python tools/time_full.py 4096 800 5
python tools/time_full.py 4096 800 5
python tools/prof_selfplay.py 1 800 300 bf16 2>&1 | grep -E "fused_tail "
