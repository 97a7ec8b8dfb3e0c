mkdir -p gpurun_out
set -x
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -k "attention" 2>&1 | tail -5
timeout 300 python profiles/kernel_bench.py attention --out gpurun_out/kb_attn_b.jsonl 2>&1 | tail -12
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"dconv_attention_tma" -c 8 -o gpurun_out/prof_attn_v2b python profiles/kernel_bench.py attention --quick > gpurun_out/ncu_attn2.log 2>&1
