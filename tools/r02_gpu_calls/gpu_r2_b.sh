#!/bin/bash
# round 2, call B: lane-group (16 lanes per game) build: GPU test-suite, 32-lane cross-check, bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2b_tests.log
tail -25 gpurun_out/r2b_tests.log
CB200_LANES=32 timeout 600 python -m pytest tests/test_gpu_trainer.py tests/test_gpu_net.py -m gpu -x -q > gpurun_out/r2b_tests32.log 2>&1
echo "tests32 rc=$?" >> gpurun_out/r2b_tests32.log
tail -5 gpurun_out/r2b_tests32.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
echo "bench rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/r2b_bench.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['kernel_ms'])"
CB200_LANES=32 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_bench32.json 2> gpurun_out/r2b_bench32.err
python -c "
import json;d=json.load(open('gpurun_out/r2b_bench32.json'));print('lanes32',d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['kernel_ms'])"
