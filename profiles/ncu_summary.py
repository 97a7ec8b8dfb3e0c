"""Summarise an .ncu-rep (ncu --set full) into one line per launch: python profiles/ncu_summary.py file.ncu-rep [--stalls]"""
import csv
import io
import subprocess
import sys

COLS = [("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "rdMB"), ("dram__bytes_write.sum", "wrMB"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"),
        ("smsp__issue_active.avg.pct", "issue%"), ("launch__registers_per_thread", "regs"),
        ("smsp__inst_executed.sum", "winst"), ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "bankconf")]
STALLS = ["barrier", "long_scoreboard", "short_scoreboard", "math_pipe_throttle", "mio_throttle", "wait", "not_selected",
          "lg_throttle", "tex_throttle", "dispatch_stall", "no_instruction", "sleeping", "membar", "branch_resolving",
          "drain", "imc_miss", "selected"]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    print(" | ".join(["kernel", "grid", "block"] + [c[1] for c in COLS]))
    for r in data:
        vals = []
        for name, short in COLS:
            if name in ix:
                v, u = r[ix[name]], units[ix[name]]
                try:
                    f = float(v.replace(",", ""))
                    if u == "byte":
                        f /= 1e6
                    elif u == "Kbyte":
                        f /= 1e3
                    elif u == "Gbyte":
                        f *= 1e3
                    elif u == "ns":
                        f /= 1e3
                    elif u == "ms":
                        f *= 1e3
                    v = f"{f:.1f}" if abs(f) < 1e6 else f"{f:.3g}"
                except ValueError:
                    pass
                vals.append(v)
            else:
                vals.append("-")
        print(" | ".join([r[ix["Kernel Name"]][:48], r[ix["Grid Size"]], r[ix["Block Size"]]] + vals))
        if "--stalls" in sys.argv:
            st = []
            for s in STALLS:
                k = f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio"
                if k in ix:
                    st.append(f"{s}={float(r[ix[k]]):.2f}")
            print("      stalls/issue: " + " ".join(st))


if __name__ == "__main__":
    main()
