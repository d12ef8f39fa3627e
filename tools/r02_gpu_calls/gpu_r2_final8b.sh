#!/bin/bash
# final defaults (6 CTAs per SM, 8 groups from 16 384 games) on 8 GPUs: configs[2], configs[3], configs[4]
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544"
$TR bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r2y_bench_n8.json 2> gpurun_out/r2y_bench_n8.err; echo "bench n8 rc=$?"
$TR tools/match_bench.py --games 10000 2> gpurun_out/r2y_match_n8.err | grep '^{' > gpurun_out/r2y_match_n8.json; echo "match n8 rc=$?"
$TR bench.py --gpus 8 --steps 2 --warmup 3 --games-per-gpu 32768 > gpurun_out/r2y_config3_n8.json 2> gpurun_out/r2y_config3_n8.err; echo "config3 n8 rc=$?"
python - <<'PY'
import json
for f in ("r2y_bench_n8","r2y_config3_n8"):
    try:
        d=json.load(open("gpurun_out/%s.json"%f)); print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["seconds_per_step"], d["e2e"]["nccl_gather_seconds_per_step"])
    except Exception as e: print(f, "ERR", e)
try: print(open("gpurun_out/r2y_match_n8.json").read()[:600])
except Exception as e: print(e)
PY
