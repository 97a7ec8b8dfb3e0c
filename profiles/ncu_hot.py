"""Hot SASS lines of one launch: python profiles/ncu_hot.py rep.ncu-rep <launch-index-1-based> [top]"""
import csv, io, subprocess, sys
rep, kid = sys.argv[1], sys.argv[2]
top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", f":::{kid}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r)
hdr = rows[hi]
data = [r for r in rows[hi + 1:] if len(r) == len(hdr) and r[hdr.index("# Samples")].isdigit()]
ia, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
tot = sum(int(r[isamp]) for r in data); totex = sum(int(r[iex]) for r in data)
print(rows[0][1][:100] if rows[0] else "", "| samples", tot, "| warp instrs", totex, "| sass lines", len(data))
for i, r in sorted(enumerate(data), key=lambda t: -int(t[1][isamp]))[:top_n]:
    print(f"{i:5d} {int(r[isamp]):6d} ({100*int(r[isamp])/max(tot,1):4.1f}%) exec {r[iex]:>9s}  {r[ia][:100]}")
B = max(50, len(data) // 24)
for b in range(0, len(data), B):
    s = sum(int(r[isamp]) for r in data[b:b + B]); e = sum(int(r[iex]) for r in data[b:b + B])
    print(f"[{b:5d}] samples {100*s/max(tot,1):5.1f}%  exec {100*e/max(totex,1):5.1f}%   {data[b][ia][:60]}")
