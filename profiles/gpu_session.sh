mkdir -p gpurun_out
s=$(date +%s)
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "tests wall $(( $(date +%s)-s )) s"
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
s=$(date +%s)
timeout 900 python bench.py > gpurun_out/v14_tedexp256.json 2> gpurun_out/v14_tedexp256.err
echo "bench wall $(( $(date +%s)-s )) s; stdout lines: $(wc -l < gpurun_out/v14_tedexp256.json)"; tail -c 300 gpurun_out/v14_tedexp256.err
timeout 600 python bench.py --workload beat-ours --no-cpu-baseline > gpurun_out/v14_beat1024.json 2> gpurun_out/v14_beat1024.err
python - <<'PY'
import json
for n in ("tedexp256", "beat1024"):
    d = json.load(open(f"gpurun_out/v14_{n}.json"))
    print(n, d["value"], d.get("ms_per_denoise_step"), d["e2e"]["value"], d["config"], d["clocks"])
PY
