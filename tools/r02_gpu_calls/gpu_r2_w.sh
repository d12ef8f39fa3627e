#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2w_bench_n1.json 2> gpurun_out/r2w_bench_n1.err; echo "bench rc=$?"; tail -n 3 gpurun_out/r2w_bench_n1.err
python -c "
import json
d=json.load(open('gpurun_out/r2w_bench_n1.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'])
print(json.dumps(d['roofline']['phases_of_the_timed_pass'],indent=1))"
