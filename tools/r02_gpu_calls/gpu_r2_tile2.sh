#!/bin/bash
# bf16x3 in the persistent kernel (8 games per CTA) + single-tile network CTA with its own shared-memory size
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python tools/time_full.py 4096 800 3 bf16x3
timeout 300 python tools/time_full.py 4096 800 6 bf16
timeout 300 python tools/time_full.py 4096 800 3 fp16
