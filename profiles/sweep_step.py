"""BASELINE config 4: single denoiser-step micro-benchmark at beat-ours dims, batch x frames sweep.

    python profiles/sweep_step.py [--out gpurun_out/sweep.jsonl] [--batches 16,64,256,1024,4096] [--frames 40,80,160]
    python profiles/sweep_step.py --one 1024x40          # three eager steps of one point (the ncu target)

Per point: one eager denoise step with a CUDA event between consecutive launches (bench.kernel_breakdown), i.e. device
time per kernel class with its algorithmic FLOPs and bytes: GEMM TFLOP/s against the measured bf16 peak, attention and
LayerNorm GB/s against the measured HBM peak.  Speech length scales with the window (800 samples per frame), so the
memory has frames*0.8 - 1 tokens (31 / 63 / 127).
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th  # noqa: E402
import gesture_b200  # noqa: E402,F401
import bench  # noqa: E402
from gesture_b200.engine import chain_for  # noqa: E402
from gesture_b200.model_creation import create_model  # noqa: E402
from gesture_b200.presets import preset  # noqa: E402
from gesture_b200.synthetic import synthetic_wav  # noqa: E402


def build_chain(model, diffusion, C, batch, frames):
    chain = chain_for(model, diffusion, (batch, C, frames), "ddpm", "cuda", use_graph=False)
    x_T = th.randn(batch, C, frames, device="cuda")
    chain.begin(x_T, synthetic_wav(batch, 800 * frames).cuda(), need_tape=False)
    return chain


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--batches", default="16,64,256,1024,4096")
    ap.add_argument("--frames", default="40,80,160")
    ap.add_argument("--one", default=None, help="BATCHxFRAMES: run three eager steps of this point only (ncu target)")
    args = ap.parse_args()
    params, C, _, _ = preset("beat-ours")
    th.manual_seed(0)
    model, diffusion, *_ = create_model(C, params)
    model.eval().to("cuda")
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(bench.ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    if args.one:
        b, f = (int(v) for v in args.one.split("x"))
        chain = build_chain(model, diffusion, C, b, f)
        for _ in range(3):
            chain.step_eager()
        th.cuda.synchronize()
        print("ok", len(chain.plan), "kernels per step")
        return
    for f in (int(v) for v in args.frames.split(",")):
        for b in (int(v) for v in args.batches.split(",")):
            chain = build_chain(model, diffusion, C, b, f)
            agg = bench.kernel_breakdown(chain)
            step_ms = sum(a["ms"] for a in agg.values())
            row = {"batch": b, "frames": f, "memory_tokens": chain.Tm, "rows": b * f, "ms_per_step": round(step_ms, 4),
                   "frames_per_s_1000_steps": round(b * f / step_ms, 1)}
            for k, a in agg.items():
                if not a["ms"]:
                    continue
                row[k] = {"ms": round(a["ms"], 4), "launches": round(a["launches"]),
                          "tflops": round(a["flops"] / (a["ms"] * 1e-3) / 1e12, 2),
                          "gbs": round(a["bytes"] / (a["ms"] * 1e-3) / 1e9, 1)}
            g = agg["gemm"]
            row["gemm_frac_of_bf16_peak"] = round(g["flops"] / (g["ms"] * 1e-3) / 1e12 / peaks.get("bf16_tflops_sustained", 1396.9), 3)
            row["attention_frac_of_hbm_peak"] = round(agg["attention"]["bytes"] / (agg["attention"]["ms"] * 1e-3) / 1e9 /
                                                      peaks.get("hbm_gbs", 6545.6), 3)
            print(json.dumps(row))
            if args.out:
                with open(args.out, "a") as fh:
                    fh.write(json.dumps(row) + "\n")
            del chain
            th.cuda.empty_cache()


if __name__ == "__main__":
    main()
