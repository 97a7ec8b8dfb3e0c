"""Summarise an ncu launch list (--metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active...,dram__... --csv):
python profiles/launch_summary.py launches.csv [first_launch last_launch]  -> per kernel class shares + per-shape table."""
import csv
import sys
from collections import OrderedDict, defaultdict

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
launch = OrderedDict()
for r in rows[hi + 1:]:
    if len(r) != len(hdr) or not r[0].isdigit():
        continue
    d = launch.setdefault(int(r[0]), {"name": r[ix["Kernel Name"]], "grid": r[ix["Grid Size"]], "block": r[ix["Block Size"]]})
    v, u = float(r[ix["Metric Value"]].replace(",", "")), r[ix["Metric Unit"]]
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
    d[r[ix["Metric Name"]]] = v * scale
ids = list(launch)
pos = [a for a in sys.argv[2:] if not a.startswith("--")]
lo = int(pos[0]) if len(pos) > 0 else ids[0]
hi_ = int(pos[1]) if len(pos) > 1 else ids[-1]
sel = [launch[i] for i in ids if lo <= i <= hi_]
T = "gpu__time_duration.sum"
TP = "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"
DR = "dram__throughput.avg.pct_of_peak_sustained_elapsed"


def cls(n):
    for k, c in (("gemm_resid_ln", "gemm_ln"), ("Li1E", "gemm_ddpm"), ("gemm_bf16", "gemm"), ("dconv_attention", "attention"),
                 ("layernorm", "layernorm"), ("scatter", "scatter"), ("step_add", "step"), ("ddpm_update", "ddpm")):
        if k in n:
            return c
    return "other"


tot = sum(d[T] for d in sel)
agg = defaultdict(lambda: [0.0, 0, 0.0, 0.0])
for d in sel:
    a = agg[cls(d["name"] if "Li1E" not in d["name"] else d["name"])]
    a[0] += d[T]; a[1] += 1; a[2] += d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0); a[3] += d.get(TP, 0) * d[T]
print(f"launches {lo}..{hi_}: {len(sel)} kernels, {tot:.1f} us (cold-cache, serialised by ncu)")
print("class | launches | us | share | dram MB | time-weighted tensor-pipe %")
for k, a in sorted(agg.items(), key=lambda t: -t[1][0]):
    print(f"{k} | {a[1]} | {a[0]:.1f} | {100 * a[0] / tot:.1f}% | {a[2]:.0f} | {a[3] / a[0]:.1f}")
if "--all" in sys.argv:
    print("\nid | kernel | grid | us | tensor % | dram % | dram rd MB | dram wr MB")
    for i in ids:
        if lo <= i <= hi_:
            d = launch[i]
            print(f"{i} | {d['name'][:58]} | {d['grid']} | {d[T]:.1f} | {d.get(TP, 0):.1f} | {d.get(DR, 0):.1f} | "
                  f"{d.get('dram__bytes_read.sum', 0):.1f} | {d.get('dram__bytes_write.sum', 0):.1f}")
