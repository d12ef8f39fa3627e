#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2g_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2g_tests.log
tail -15 gpurun_out/r2g_tests.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/r2g_bench.err
python -c "
import json;d=json.load(open('gpurun_out/r2g_bench.json'));print(d['value'],d['ms_per_step'],d['e2e'])"
