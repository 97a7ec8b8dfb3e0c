mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_chain_gpu.py -x -q -k "speech_encoder_implementations or eval_infer_time or batch_invariance" 2>&1 | tail -8
timeout 300 python bench.py --workload beat-ours --no-cpu-baseline --steps 1 --warmup 3 > gpurun_out/v16b_beat1024.json 2> gpurun_out/v16b.err; python -c "
import json; d=json.load(open('gpurun_out/v16b_beat1024.json')); print(d['value'], d['chain_begin_ms'], d['speech_encoder'])"
