timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest4.log
tail -4 gpurun_out/pytest4.log
for c in 64 128 256 1024; do GD_FUSE_LN=1 python profiles/ingraph_breakdown.py --workload beat-ours --clips $c > gpurun_out/ingraph_fuse_beat$c.json 2> gpurun_out/ingraph_fuse_beat$c.err; python profiles/ingraph_breakdown.py --workload beat-ours --clips $c > gpurun_out/ingraph4_beat$c.json 2>/dev/null; done
python profiles/ingraph_breakdown.py --workload tedexp-ours --clips 256 > gpurun_out/ingraph4_tedexp256.json 2>/dev/null
python bench.py --steps 2 --warmup 3 > gpurun_out/bench_r02_b.json 2> gpurun_out/bench_r02_b.err; echo bench rc=$?
