#!/bin/bash
# HEAD on 2 GPUs: the driver's scaling command + the reference arm under torchrun (rank 0 alone works)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29546"
$TR bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2zz_bench_n2.json 2> gpurun_out/r2zz_bench_n2.err; echo "bench n2 rc=$?"
tail -n 2 gpurun_out/r2zz_bench_n2.err
$TR bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2zz_ref_n2.json 2> gpurun_out/r2zz_ref_n2.err; echo "ref n2 rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2zz_bench_n2.json")); print("n2", d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"].get("nccl_gather_seconds_per_step"))
r=json.load(open("gpurun_out/r2zz_ref_n2.json")); print("ref", r["value"], r.get("impl"))
PY
