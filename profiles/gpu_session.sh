mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -k "linear" 2>&1 | tail -3
timeout 300 python profiles/kernel_bench.py gemm --out gpurun_out/kb_gemm_e64.jsonl 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['M'], d['N'], d['K'], d['mode'], d['us'], d['TFLOPs'])"
timeout 300 python profiles/cublas_ref.py 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('cublas', d['M'], d['N'], d['K'], d['us'], d['TFLOPs'])"
for w in tedexp-ours beat-ours; do timeout 300 python bench.py --workload $w --steps 1 --warmup 1 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$w', round(d['value'],1), round(d['ms_per_denoise_step'],3), {k:v['ms_per_step'] for k,v in d['kernel_breakdown'].items()}, d['clocks']['sm_mhz'])"; done
