"""A/B of SplitChain (engine.py): ms per denoise step of a whole-chain graph with the rank's clips sampled as 1, 2, 4 (8)
parallel sub-chains.  One JSON line per (workload, clips, parts).

    python profiles/subchain_sweep.py > profiles/r02_subchain_sweep.jsonl
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch as th  # noqa: E402


def main():
    import gesture_b200  # noqa: F401
    from gesture_b200.engine import chain_for, release_chains
    from gesture_b200.model_creation import create_model
    from gesture_b200.synthetic import synthetic_wav
    from bench import workload_preset
    cases = [("beat-ours", c, p) for c in (64, 128, 256, 512, 1024) for p in (1, 2, 4, 8) if c // p >= 16] + \
            [("tedexp-ours", c, p) for c in (32, 64, 128, 256) for p in (1, 2, 4) if c // p >= 16] + \
            [("beat-ours-4x", c, p) for c in (64, 512) for p in (1, 2, 4)]
    if len(sys.argv) > 1:
        cases = [c for c in cases if c[0] in sys.argv[1:]]
    last = None
    for wl, clips, parts in cases:
        params, C, T, L, _ = workload_preset(wl)
        params["Diffusion"]["timestep_respacing"] = "ddim200"  # 200-step process: same kernels per step, shorter capture
        if last != wl:
            th.manual_seed(0)
            model, diffusion, *_ = create_model(C, params)
            model.eval().to("cuda")
            last = wl
        model.sub_chains = parts
        shape = (clips, C, T)
        chain = chain_for(model, diffusion, shape, "ddpm", "cuda", allow_split=True)
        x_T = th.randn(shape, device="cuda")
        wav = synthetic_wav(clips, L, seed=1).cuda()
        chain.begin(x_T, wav)
        chain.run()
        th.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            chain.begin(x_T, wav)
            th.cuda.synchronize()
            e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
            e0.record()
            chain.run()
            e1.record()
            th.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        n = diffusion.num_timesteps
        print(json.dumps({"workload": wl, "clips": clips, "parts": getattr(chain, "parts", 1), "us_per_denoise_step": round(best * 1e3 / n, 1),
                          "frames_per_s": round(clips * T * 1000 / n / (best / 1e3) / 1000 * n, 0) if False else round(clips * T / (best / 1e3 / n * 1000), 0),
                          "graph": chain.graph_info}), flush=True)
        release_chains(model)


if __name__ == "__main__":
    main()
