mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_speech_gpu.py -x -q 2>&1 | tail -15
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_speech_gpu.py 2>&1 | tail -6
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 300 python profiles/begin_breakdown.py --workload beat-ours 2>&1 | tail -1
GD_SPEECH=native-bf16 timeout 300 python profiles/begin_breakdown.py --workload beat-ours 2>&1 | tail -1
