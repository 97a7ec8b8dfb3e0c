mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v13_tedexp256_2gpu.json 2> gpurun_out/bench_2gpu.err; tail -c 300 gpurun_out/bench_2gpu.err; python -c "
import json; d=json.load(open('gpurun_out/bench_v13_tedexp256_2gpu.json')); print(d['n_gpus'], d['value'], d['ms_per_denoise_step'], d['e2e']['value'], d['clocks'])"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --impl reference --steps 2 --warmup 1 2>/dev/null | cut -c1-400
timeout 300 python -m pytest tests/test_chain_gpu.py -x -q 2>&1 | tail -2
