#!/bin/bash
mkdir -p gpurun_out
python tools/san_small.py > gpurun_out/r2t_plain.log 2>&1; tail -2 gpurun_out/r2t_plain.log
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 3 python tools/san_small.py > gpurun_out/r2t_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -4 gpurun_out/r2t_memcheck.log
