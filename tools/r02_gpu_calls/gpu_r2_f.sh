#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/r2f_bench.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2f_ref.json 2> gpurun_out/r2f_ref.err
echo "ref rc=$?"; tail -3 gpurun_out/r2f_ref.err
