"""Per-class and per-launch summary of an ncu launch list of the speech encoder (second pass of
`begin_breakdown.py --encoder-only N`):  python profiles/encoder_launch_summary.py launches.csv > summary.txt"""
import collections
import csv
import re
import sys


def main():
    with open(sys.argv[1]) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    recs = collections.OrderedDict()
    for row in csv.DictReader(lines):
        r = recs.setdefault(row["ID"], {"name": row["Kernel Name"], "grid": row["Grid Size"]})
        try:
            r[row["Metric Name"]] = float(row["Metric Value"].replace(",", ""))
        except ValueError:
            pass
    ids = list(recs)
    sel = [recs[i] for i in ids[len(ids) // 2:]]  # the second (warm) pass
    T = "gpu__time_duration.sum"
    tensor = "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active"
    total = sum(x.get(T, 0) for x in sel) / 1e3
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
    for x in sel:
        nm = x["name"].split("(")[0]
        m = re.search(r"gemm_bf16_tn_kernel<(\d+), (\d+), (\d+)>", x["name"])
        if m:
            nm = "gemm<BN=%s,mode=%s>%s" % (m.group(1), m.group(2), " (MODE_CONV)" if m.group(2) == "4" else "")
        a = agg[nm[:44]]
        a[0] += 1
        a[1] += x.get(T, 0) / 1e3
        a[2] += x.get("dram__bytes_read.sum", 0) + x.get("dram__bytes_write.sum", 0)
    print(f"launches {len(sel)}, {total:.1f} us (cold-cache, serialised by ncu)")
    print("class | launches | us | share | dram MB")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k} | {v[0]} | {v[1]:.1f} | {v[1] / total:.1%} | {v[2] / 1e6:.1f}")
    print("\n# | kernel | grid | us | tensor % | dram MB")
    for i, x in enumerate(sel):
        mb = (x.get("dram__bytes_read.sum", 0) + x.get("dram__bytes_write.sum", 0)) / 1e6
        print(f"{i} | {x['name'][:58]} | {x['grid']} | {x.get(T, 0) / 1e3:.1f} | {x.get(tensor, 0):.1f} | {mb:.1f}")


if __name__ == "__main__":
    main()
