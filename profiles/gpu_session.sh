mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_speech_gpu.py -x -q 2>&1 | tail -12
timeout 300 python profiles/begin_breakdown.py --workload beat-ours 2>&1 | tail -1
