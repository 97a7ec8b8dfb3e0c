mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -k "layernorm or ddpm" 2>&1 | tail -3
timeout 300 python profiles/kernel_bench.py layernorm 2>&1 | tail -4
timeout 300 python profiles/kernel_bench.py ddpm 2>&1 | tail -4
for w in tedexp-ours beat-ours; do timeout 300 python bench.py --workload $w --steps 1 --warmup 1 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$w', round(d['value'],1), round(d['ms_per_denoise_step'],3), {k:v['ms_per_step'] for k,v in d['kernel_breakdown'].items()}, d['clocks']['sm_mhz'])"; done
