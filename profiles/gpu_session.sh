mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu -s 2>&1 | grep -E "inpaint-model|passed|failed|Error|error" | tail -12
