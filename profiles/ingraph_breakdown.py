"""In-graph per-kernel-class times: CUPTI activity records (torch.profiler / Kineto) of ONE
replay of a K-step CUDA graph of the sampling chain - the kernels as they really run inside the graph (programmatic dependent
launch, concurrent branches), not the eager step with events between launches that bench.py's `kernel_breakdown` times.

    python profiles/ingraph_breakdown.py --workload tedexp-ours --clips 256 [--graph-steps 10] > profiles/r02_ingraph_breakdown_tedexp256.json

Prints one JSON object: per class {launches/step, busy us/step (sum of kernel durations), share}, the wall time per step of
the profiled replay (first kernel start -> last kernel end), the un-profiled replay time next to it (profiling overhead),
and the idle time per step during which NO kernel of the graph was running (launch gaps).
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch as th  # noqa: E402


def classify(name):
    if "gemm_ln_a" in name:
        return "gemm_ln_prologue"
    if "gemm_resid_ln" in name:
        return "gemm_ln"
    if "gemm_bf16_tn_kernel" in name:
        m = name.split("gemm_bf16_tn_kernel<")[1].split(">")[0].replace(" ", "").split(",")
        mode = {"0": "direct", "1": "ddpm", "2": "bf16out", "3": "f32_reduce_add"}.get(m[1], m[1])
        return f"gemm[{mode}]"
    for key, cls in (("dconv_attention", "attention"), ("layernorm", "layernorm"), ("scatter_row", "scatter"), ("step_add", "step")):
        if key in name:
            return cls
    return "other:" + name[:40]


def measure(model, diffusion, shape, wav, x_T, graph_steps=10):
    """-> dict: CUPTI records of one replay of a `graph_steps`-step graph of the chain (model.graph_steps is restored)."""
    from gesture_b200.engine import chain_for
    saved = getattr(model, "graph_steps", 0)
    model.graph_steps = graph_steps
    try:
        chain = chain_for(model, diffusion, shape, "ddpm", x_T.device)
        chain.begin(x_T, wav)
        chain.run(n_steps=3 * graph_steps)  # capture + warm replays
        th.cuda.synchronize()
        e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
        e0.record()
        chain.graph.replay()
        e1.record()
        th.cuda.synchronize()
        plain_us = e0.elapsed_time(e1) * 1e3 / graph_steps
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            chain.graph.replay()
            th.cuda.synchronize()
    finally:
        model.graph_steps = saved
    evs = [e for e in prof.events() if e.device_type == th.autograd.DeviceType.CUDA and "Memcpy" not in e.name and "Memset" not in e.name]
    evs.sort(key=lambda e: e.time_range.start)
    K = graph_steps
    agg = {}
    for e in evs:
        c = agg.setdefault(classify(e.name), {"launches": 0, "us": 0.0})
        c["launches"] += 1
        c["us"] += e.time_range.end - e.time_range.start
    t0, t1 = evs[0].time_range.start, max(e.time_range.end for e in evs)
    busy, cur_s, cur_e = 0.0, None, None  # union of busy intervals -> time with no kernel running
    for e in evs:
        s, t = e.time_range.start, e.time_range.end
        if cur_e is None or s > cur_e:
            if cur_e is not None:
                busy += cur_e - cur_s
            cur_s, cur_e = s, t
        else:
            cur_e = max(cur_e, t)
    busy += cur_e - cur_s
    total = sum(c["us"] for c in agg.values())
    return {"graph_steps": K, "kernels_per_step": len(evs) / K, "wall_us_per_step_profiled": (t1 - t0) / K,
            "wall_us_per_step_unprofiled": plain_us, "idle_us_per_step": ((t1 - t0) - busy) / K,
            "sum_kernel_us_per_step": total / K,
            "note": "CUPTI kernel records of one graph replay; under programmatic dependent launch a kernel's duration includes the "
                    "time it waits for its predecessor, so the class sums over-count and wall < sum",
            "classes": {k: {"launches_per_step": v["launches"] / K, "us_per_step": round(v["us"] / K, 2),
                            "avg_us": round(v["us"] / v["launches"], 2), "share_of_kernel_time": round(v["us"] / total, 4)}
                        for k, v in sorted(agg.items(), key=lambda kv: -kv[1]["us"])}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="tedexp-ours")
    ap.add_argument("--clips", type=int, default=256)
    ap.add_argument("--graph-steps", type=int, default=10)
    a = ap.parse_args()
    import gesture_b200  # noqa: F401
    from gesture_b200.model_creation import create_model
    from gesture_b200.synthetic import synthetic_wav
    from bench import workload_preset
    params, C, T, L, _ = workload_preset(a.workload)
    th.manual_seed(0)
    model, diffusion, *_ = create_model(C, params)
    model.eval().to("cuda")
    shape = (a.clips, C, T)
    out = {"workload": a.workload, "clips": a.clips}
    out.update(measure(model, diffusion, shape, synthetic_wav(a.clips, L, seed=1).cuda(), th.randn(shape, device="cuda"), a.graph_steps))
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
