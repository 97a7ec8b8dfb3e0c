mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_chain_gpu.py -x -q 2>&1 | tail -3
for cfg in "1 0" "0 0" "1 1" "0 1"; do set -- $cfg
GD_PDL=$1 GD_FUSE_LN=$2 timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('tedexp pdl=$1 fuse_ln=$2', round(d['value'],1), round(d['ms_per_denoise_step'],3), d['clocks']['sm_mhz'])"
done
for cfg in "1 0" "0 0" "1 1"; do set -- $cfg
GD_PDL=$1 GD_FUSE_LN=$2 timeout 300 python bench.py --workload beat-ours --steps 1 --warmup 1 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('beat pdl=$1 fuse_ln=$2', round(d['value'],1), round(d['ms_per_denoise_step'],3), d['clocks']['sm_mhz'])"
done
