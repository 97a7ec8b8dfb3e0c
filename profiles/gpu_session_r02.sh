# Round-2 validation session (one `gpurun -- 'bash profiles/gpu_session_r02.sh'` call): GPU parity tests, smoke, ncu launch
# list of one tedexp / beat denoise step, one `ncu --set full` capture of the dominant GEMM, the attention kernel and the
# LayerNorm-prologue GEMM.  Outputs land in gpurun_out/; the summaries are copied to profiles/r02_*.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu.log; tail -4 gpurun_out/r02_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__throughput.avg.pct_of_peak_sustained_elapsed
python profiles/profile_step.py tedexp-ours 256 > gpurun_out/plain_tedexp.log 2>&1 && \
ncu --metrics $M --clock-control none -k 'regex:gemm_|dconv_attention|layernorm|scatter_row|step_add' --csv --log-file gpurun_out/r02_launches_tedexp256.csv python profiles/profile_step.py tedexp-ours 256 > gpurun_out/ncu_tedexp.log 2>&1
python profiles/profile_step.py beat-ours 1024 > gpurun_out/plain_beat.log 2>&1 && \
ncu --metrics $M --clock-control none -k 'regex:gemm_|dconv_attention|layernorm|scatter_row|step_add' --csv --log-file gpurun_out/r02_launches_beat1024.csv python profiles/profile_step.py beat-ours 1024 > gpurun_out/ncu_beat.log 2>&1
# full captures (one step = 192 launches; skip the first step)
ncu --set full --clock-control none --import-source on -k 'regex:gemm_bf16_tn_kernel' -s 120 -c 12 -o gpurun_out/r02_gemm_full python profiles/profile_step.py tedexp-ours 256 > gpurun_out/ncu_full_gemm.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:dconv_attention' -s 33 -c 4 -o gpurun_out/r02_attn_full python profiles/profile_step.py tedexp-ours 256 > gpurun_out/ncu_full_attn.log 2>&1
GD_LN_PROLOGUE=1 ncu --set full --clock-control none --import-source on -k 'regex:gemm_ln_a' -s 52 -c 5 -o gpurun_out/r02_lnprologue_full python profiles/profile_step.py tedexp-ours 256 > gpurun_out/ncu_full_lnp.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/r02_launches_*.csv
