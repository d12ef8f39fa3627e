#!/bin/bash
# single-tile network CTA with its own shared-memory size (77 KB instead of 105 KB; 104 registers) and a bf16x3
# form of it (154 KB), so that bf16x3 runs overlap network and game step like the other modes
timeout 900 python -m pytest tests/test_gpu_net.py tests/test_gpu_bench_config.py tests/test_gpu_tourney.py -x -q 2>&1 | tail -3
timeout 300 python tools/time_full.py 4096 800 6 bf16
timeout 300 python tools/time_full.py 4096 800 6 bf16
timeout 300 python tools/time_full.py 4096 800 3 bf16x3
timeout 300 python tools/time_full.py 4096 800 3 fp16
timeout 600 python tools/time_full.py 32768 800 2 bf16
