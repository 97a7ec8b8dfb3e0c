timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -k "layernorm or linear" > gpurun_out/pytest5.log 2>&1; echo "pytest kernels rc=$?" >> gpurun_out/pytest5.log
timeout 600 python -m pytest tests/test_chain_gpu.py -x -q -k "prologue or p_sample_dict" >> gpurun_out/pytest5.log 2>&1; echo "pytest chain rc=$?" >> gpurun_out/pytest5.log
tail -15 gpurun_out/pytest5.log
for c in 128 1024; do GD_LN_PROLOGUE=1 timeout 300 python profiles/ingraph_breakdown.py --workload beat-ours --clips $c > gpurun_out/ingraph_lnp_beat$c.json 2> gpurun_out/ingraph_lnp_beat$c.err; done
GD_LN_PROLOGUE=1 timeout 300 python profiles/ingraph_breakdown.py --workload tedexp-ours --clips 256 > gpurun_out/ingraph_lnp_tedexp256.json 2> gpurun_out/ingraph_lnp_tedexp256.err
timeout 300 python profiles/ingraph_breakdown.py --workload tedexp-ours --clips 256 > gpurun_out/ingraph5_tedexp256.json 2> /dev/null
grep -h "wall_us_per_step_unprofiled" gpurun_out/ingraph_lnp_*.json gpurun_out/ingraph5_tedexp256.json
