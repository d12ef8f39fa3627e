#!/bin/bash
# final build: the driver's own sequence -- GPU tests, smoke, bench (engine + reference arm)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2z_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2z_tests.log; tail -3 gpurun_out/r2z_tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2z_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2z_smoke.log
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2z_bench_ref.json 2> gpurun_out/r2z_bench_ref.err; echo "ref rc=$?"
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2z_bench_n1.json 2> gpurun_out/r2z_bench_n1.err; echo "bench rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/r2z_bench_n1.json')); r=json.load(open('gpurun_out/r2z_bench_ref.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'ref',r['value'],'ratio e2e',d['e2e']['value']/r['value'])
print('roofline',d['roofline']['frac'],d['roofline']['kernel_time_share'],'K1',d['game_logic']['states_per_sec'],d['game_logic']['roofline']['frac'])
print('cpu',d['cpu_baseline'])"
