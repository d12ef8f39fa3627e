timeout 600 python -m pytest tests/test_gpu_net.py -q -m gpu --timeout 200 -s -k "trained" 2>&1 | grep -E "passed|failed|trained net|trained vs|Error" | head -12
