mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_chain_gpu.py -x -q 2>&1 | tail -3
