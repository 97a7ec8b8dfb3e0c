# final check of the round-2 build: GPU tests, smoke, N=1 bench with the driver's flags
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r02_final_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_final_pytest.log; tail -3 gpurun_out/r02_final_pytest.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/r02_final_bench_n1.json 2> gpurun_out/r02_final_bench_n1.err; echo bench rc=$?
