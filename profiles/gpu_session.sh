mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/bench_v13_tedexp256.json 2> gpurun_out/bench_v13.err; tail -c 200 gpurun_out/bench_v13.err; wc -l gpurun_out/bench_v13_tedexp256.json; python -c "
import json; d=json.load(open('gpurun_out/bench_v13_tedexp256.json')); print(d['value'], d['ms_per_denoise_step'], d['e2e']['value'], d['roofline']['achieved'], d['clocks'], {k:v['ms_per_step'] for k,v in d['kernel_breakdown'].items()})"
timeout 600 python bench.py --workload beat-ours > gpurun_out/bench_v13_beat1024.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_v13_beat1024.json')); print(d['value'], d['ms_per_denoise_step'], d['e2e']['value'], d['roofline']['achieved'], d['clocks'], {k:v['ms_per_step'] for k,v in d['kernel_breakdown'].items()})"
timeout 600 python bench.py --workload beat-ours-4x --no-cpu-baseline > gpurun_out/bench_v13_beat4x64.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_v13_beat4x64.json')); print(d['value'], d['ms_per_denoise_step'])"
