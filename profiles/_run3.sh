for c in 256; do python profiles/ingraph_breakdown.py --workload tedexp-ours --clips $c > gpurun_out/ingraph_tedexp$c.json 2> gpurun_out/ingraph_tedexp$c.err; done
for c in 64 128 256 1024; do python profiles/ingraph_breakdown.py --workload beat-ours --clips $c > gpurun_out/ingraph_beat$c.json 2> gpurun_out/ingraph_beat$c.err; done
tail -3 gpurun_out/ingraph_beat128.err; head -c 1500 gpurun_out/ingraph_beat128.json
