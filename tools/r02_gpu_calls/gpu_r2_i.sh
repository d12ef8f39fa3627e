#!/bin/bash
# N=2 smoke of every multi-rank command before the 8-GPU call
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
$TR bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/r2i_bench_n2.json 2> gpurun_out/r2i_bench_n2.err; echo "bench n2 rc=$?"
$TR tools/match_bench.py --games 2500 > gpurun_out/r2i_match_n2.json 2> gpurun_out/r2i_match_n2.err; echo "match n2 rc=$?"
$TR bench.py --gpus 2 --steps 1 --warmup 3 --games-per-gpu 32768 > gpurun_out/r2i_config3_n2.json 2> gpurun_out/r2i_config3_n2.err; echo "config3 n2 rc=$?"
tail -2 gpurun_out/r2i_*.err
python - <<'PY'
import json
for f in ("r2i_bench_n2","r2i_config3_n2"):
    try:
        d=json.load(open("gpurun_out/%s.json"%f)); print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["seconds_per_step"], d["e2e"]["nccl_gather_seconds_per_step"])
    except Exception as e: print(f, "ERR", e)
try: print(open("gpurun_out/r2i_match_n2.json").read()[:900])
except Exception as e: print(e)
PY
