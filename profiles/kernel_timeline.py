"""Serial path of a denoise step from in-kernel %globaltimer stamps (profiling build `csrc/build.py --trace`, -DGD_TRACE).

    GD_LIB=<pkg>/libgd_b200_trace.so python profiles/kernel_timeline.py --workload beat-ours --clips 128

Block 0 of every kernel of ONE replay of a 10-step graph stamps: 0 entry, 1 prologue done, 2 predecessor complete
(griddepcontrol.wait returned), 3 first operands landed, 4 MMAs issued, 5 accumulator complete, 6 last store issued, 7 stores
drained, 8 exit.  Printed per kernel of the last captured step: how long it waited for its predecessor after its own
prologue, and where the time from "predecessor complete" to "exit" went; `release` = from this kernel's exit stamp to the next
kernel's "predecessor complete" (grid completion + dependent release).
"""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch as th  # noqa: E402

KIND = {100: "gemm[direct]", 101: "gemm[ddpm]", 102: "gemm[bf16out]", 103: "gemm[f32 reduce-add]", 200: "layernorm", 300: "scatter",
        400: "step", 500: "attention"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="beat-ours")
    ap.add_argument("--clips", type=int, default=128)
    ap.add_argument("--graph-steps", type=int, default=10)
    a = ap.parse_args()
    import gesture_b200  # noqa: F401
    from gesture_b200 import _lib
    from gesture_b200.engine import chain_for
    from gesture_b200.model_creation import create_model
    from gesture_b200.synthetic import synthetic_wav
    from bench import workload_preset
    lib = _lib.load()
    if not hasattr(lib, "gd_debug_set_trace"):
        raise SystemExit("needs the profiling build: python <pkg>/csrc/build.py --trace ; GD_LIB=<pkg>/libgd_b200_trace.so")
    params, Cc, T, L, _ = workload_preset(a.workload)
    th.manual_seed(0)
    model, diffusion, *_ = create_model(Cc, params)
    model.eval().to("cuda")
    model.graph_steps = a.graph_steps
    shape = (a.clips, Cc, T)
    chain = chain_for(model, diffusion, shape, "ddpm", "cuda")
    chain.begin(th.randn(shape, device="cuda"), synthetic_wav(a.clips, L, seed=1).cuda())
    chain.run(n_steps=3 * a.graph_steps)
    th.cuda.synchronize()
    per_step = len(chain.plan)
    cap = per_step * a.graph_steps + 64
    buf = th.zeros(16 + cap * 10, dtype=th.int64, device="cuda")
    buf[1] = cap
    lib.gd_debug_set_trace.restype, lib.gd_debug_set_trace.argtypes = C.c_int32, [C.c_void_p]
    assert lib.gd_debug_set_trace(buf.data_ptr()) == 0
    chain.graph.replay()
    th.cuda.synchronize()
    lib.gd_debug_set_trace(None)
    host = buf.cpu().numpy()
    n = int(host[0])
    slots = host[16:16 + min(n, cap) * 10].reshape(-1, 10)
    traced = len(slots)
    last = slots[-per_step:] if traced >= per_step else slots
    rows, totals = [], {}
    for i, s in enumerate(last):
        kind = KIND.get(int(s[9]), str(int(s[9])))
        nxt = last[i + 1] if i + 1 < len(last) else None
        t = [int(v) for v in s[:9]]
        start = t[2]
        r = {"kernel": kind,
             "waited_for_predecessor_us": round((t[2] - (t[1] or t[0])) / 1e3, 2) if (t[1] or t[0]) else None,
             "first_operands_us": round((t[3] - t[2]) / 1e3, 2) if t[3] else None,
             "mma_us": round((t[4] - t[3]) / 1e3, 2) if t[3] and t[4] else None,
             "accumulator_us": round((t[5] - t[4]) / 1e3, 2) if t[4] and t[5] else None,
             "epilogue_us": round((t[6] - t[5]) / 1e3, 2) if t[5] and t[6] else None,
             "drain_us": round((t[7] - t[6]) / 1e3, 2) if t[6] and t[7] else None,
             "body_us": round((t[8] - t[2]) / 1e3, 2) if t[8] else None,
             "release_us": round((int(nxt[2]) - t[8]) / 1e3, 2) if nxt is not None and t[8] and int(nxt[2]) else None,
             "period_us": round((int(nxt[2]) - start) / 1e3, 2) if nxt is not None and int(nxt[2]) else None}
        rows.append(r)
        tt = totals.setdefault(kind, {"n": 0, "period_us": 0.0, "body_us": 0.0, "release_us": 0.0})
        tt["n"] += 1
        for k in ("period_us", "body_us", "release_us"):
            tt[k] += r[k] or 0.0
    step_us = (int(last[-1][8]) - int(last[0][2])) / 1e3 if len(last) > 1 else None
    print(json.dumps({"workload": a.workload, "clips": a.clips, "kernels_traced": traced, "kernels_per_step": per_step,
                      "last_step_us": step_us, "by_class": {k: {kk: round(vv, 2) for kk, vv in v.items()} for k, v in totals.items()},
                      "kernels": rows}, indent=1))


if __name__ == "__main__":
    main()
