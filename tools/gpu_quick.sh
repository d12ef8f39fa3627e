timeout 600 python -m pytest tests/test_gpu_net.py -q -m gpu --timeout 200 -x -s 2>&1 | grep -E "passed|failed|bf16|Error|error" | head
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_n1.json')); print({k:d[k] for k in ['value','ms_per_step','ms_per_step_profiled','moves_per_sec','gpu_launches']}); print(d['e2e']); print(d['roofline']); print(d['cpu_baseline']); print(d['clocks']); print(d['game_logic'])"
