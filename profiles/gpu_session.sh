mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
timeout 600 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err; tail -c 600 gpurun_out/bench_r2a.err; cat gpurun_out/bench_r2a.json
timeout 600 python bench.py --workload beat-ours --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2a_beat.json 2> gpurun_out/bench_r2a_beat.err; cat gpurun_out/bench_r2a_beat.json
