#!/bin/bash
# closing build on 8 GPUs: configs[2] and configs[3]
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544"
$TR bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r2x_bench_n8.json 2> gpurun_out/r2x_bench_n8.err; echo "bench n8 rc=$?"
$TR bench.py --gpus 8 --steps 2 --warmup 3 --games-per-gpu 32768 > gpurun_out/r2x_config3_n8.json 2> gpurun_out/r2x_config3_n8.err; echo "config3 n8 rc=$?"
python - <<'PY'
import json
for f in ("r2x_bench_n8","r2x_config3_n8"):
    try:
        d=json.load(open("gpurun_out/%s.json"%f)); print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["seconds_per_step"], d["e2e"]["nccl_gather_seconds_per_step"])
    except Exception as e: print(f, "ERR", e)
PY
