# The end-of-round validation session (one `gpurun -- 'bash profiles/gpu_session.sh'` call, ~4 GPU-minutes):
# GPU parity tests, smoke, the three bench workloads, and the conditioning breakdown.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/v16_tedexp256.json 2> gpurun_out/v16_tedexp256.err
timeout 600 python bench.py --workload beat-ours --no-cpu-baseline > gpurun_out/v16_beat1024.json 2> gpurun_out/v16_beat1024.err
timeout 600 python bench.py --workload beat-ours-4x --no-cpu-baseline > gpurun_out/v16_beat4x64.json 2> gpurun_out/v16_beat4x64.err
python - <<'PY'
import json
for n in ("tedexp256", "beat1024", "beat4x64"):
    d = json.load(open(f"gpurun_out/v16_{n}.json"))
    print(n, round(d["value"], 1), d.get("ms_per_denoise_step"), round(d["ms_per_step"], 1), round(d["e2e"]["value"], 1), d["clocks"], d["gpu_launches"])
PY
for w in beat-ours tedexp-ours beat-ours-4x; do timeout 300 python profiles/begin_breakdown.py --workload $w 2>&1 | tail -1 | tee -a gpurun_out/begin_breakdown_v16.jsonl | cut -c1-220; done
# encoder launch list (ncu, one 64-clip pass):
#   ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active \
#       --clock-control none -k 'regex:gemm_bf16|speech_stem|se_gate|se_residual|pixel_shuffle|mel_power|instance_norm' --csv \
#       --log-file gpurun_out/enc_launches.csv python profiles/begin_breakdown.py --workload beat-ours --encoder-only 64
#   python profiles/encoder_launch_summary.py gpurun_out/enc_launches.csv
