#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_trainer.py -m gpu -x -q > gpurun_out/r2l_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2l_tests.log; tail -12 gpurun_out/r2l_tests.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
$TR bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/r2l_bench_n2.json 2> gpurun_out/r2l_bench_n2.err; echo "bench n2 rc=$?"
tail -n 3 gpurun_out/r2l_bench_n2.err
python -c "
import json;d=json.load(open('gpurun_out/r2l_bench_n2.json'));print(d['value'],d['ms_per_step'],d['e2e'])"
