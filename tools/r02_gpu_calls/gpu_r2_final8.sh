#!/bin/bash
# final HEAD on 8 GPUs: the driver's scaling command, with the C-ABI NCCL gather in the e2e leg
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544"
$TR bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r2zz_bench_n8.json 2> gpurun_out/r2zz_bench_n8.err; echo "bench n8 rc=$?"
tail -n 2 gpurun_out/r2zz_bench_n8.err
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29545"
$TR4 bench.py --gpus 4 --steps 3 --warmup 3 > gpurun_out/r2zz_bench_n4.json 2> gpurun_out/r2zz_bench_n4.err; echo "bench n4 rc=$?"
python - <<'PY'
import json
for f in ("r2zz_bench_n8","r2zz_bench_n4"):
    d=json.load(open("gpurun_out/%s.json"%f)); print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["seconds_per_step"], d["e2e"]["nccl_gather_seconds_per_step"])
PY
