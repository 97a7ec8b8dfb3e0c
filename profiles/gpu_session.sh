mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -k "attention" 2>&1 | tail -12
GD_ATTN=v3 timeout 300 python profiles/kernel_bench.py attention 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['dk'], d['rows_q'], d['rows_kv'], d['us'])
    else: print(l.rstrip()[:200])" | tail -12
GD_ATTN=v3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"dconv_attention_tc" -c 4 -o gpurun_out/prof_attn_tc python profiles/kernel_bench.py attention --quick > gpurun_out/ncu_attn_tc.log 2>&1
