mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout 300 > gpurun_out/t_gpu_full.log 2>&1; echo "gpu tests rc=$?"; tail -4 gpurun_out/t_gpu_full.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
