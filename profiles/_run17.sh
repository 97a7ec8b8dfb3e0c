timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/pytest17.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest17.log; tail -3 gpurun_out/pytest17.log
python profiles/chain_ab.py tedexp-ours 256 base base > gpurun_out/ab_lnsplit.jsonl 2>&1; cut -c 1-200 gpurun_out/ab_lnsplit.jsonl
