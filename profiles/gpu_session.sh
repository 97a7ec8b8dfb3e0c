mkdir -p gpurun_out
K='regex:gemm_bf16|speech_stem|se_gate|se_residual|pixel_shuffle|mel_power|instance_norm'
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k "$K" --csv --log-file gpurun_out/enc_launches.csv python profiles/begin_breakdown.py --workload beat-ours --encoder-only 64 > gpurun_out/enc_ncu.log 2>&1
tail -1 gpurun_out/enc_ncu.log; wc -l gpurun_out/enc_launches.csv
: > gpurun_out/enc_ncu_full_summary.txt
for skip in 41 50 61 73; do
timeout 200 ncu --set full --clock-control none -k regex:gemm_bf16 --launch-skip $skip --launch-count 1 -f -o /tmp/enc_full_$skip python profiles/begin_breakdown.py --workload beat-ours --encoder-only 64 > /tmp/enc_full_$skip.log 2>&1
python profiles/ncu_summary.py /tmp/enc_full_$skip.ncu-rep --stalls >> gpurun_out/enc_ncu_full_summary.txt 2>&1
done
cat gpurun_out/enc_ncu_full_summary.txt | cut -c1-250
