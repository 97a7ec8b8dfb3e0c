"""A/B of engine switches on the whole-chain graph: ms per denoise step of a 200-step respaced chain (same kernels per step
as the 1000-step chain), best of 3 replays, one JSON line per variant.

    python profiles/chain_ab.py tedexp-ours 256 GD_SM_PARTITION=1 GD_SM_PARTITION=0 ...
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(workload, clips):
    import torch as th
    import gesture_b200  # noqa: F401
    from gesture_b200.engine import chain_for
    from gesture_b200.model_creation import create_model
    from gesture_b200.synthetic import synthetic_wav
    from bench import workload_preset
    params, C, T, L, _ = workload_preset(workload)
    params["Diffusion"]["timestep_respacing"] = "ddim200"
    th.manual_seed(0)
    model, diffusion, *_ = create_model(C, params)
    model.eval().to("cuda")
    shape = (clips, C, T)
    chain = chain_for(model, diffusion, shape, "ddpm", "cuda", allow_split=True)
    x_T = th.randn(shape, device="cuda")
    wav = synthetic_wav(clips, L, seed=1).cuda()
    chain.begin(x_T, wav)
    out = chain.run()["sample"]
    th.cuda.synchronize()
    times = []
    for _ in range(4):
        chain.begin(x_T, wav)
        th.cuda.synchronize()
        e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
        e0.record()
        chain.run()
        e1.record()
        th.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    import hashlib
    n = diffusion.num_timesteps
    print("AB " + json.dumps({"us_per_step_best": round(min(times) * 1e3 / n, 1), "us_per_step_all": [round(t * 1e3 / n, 1) for t in times],
                              "kernels_per_step": len(chain.plan), "finite": bool(th.isfinite(out).all())}))


def main():
    if sys.argv[1] == "--child":
        child(sys.argv[2], int(sys.argv[3]))
        return
    workload, clips, variants = sys.argv[1], int(sys.argv[2]), sys.argv[3:]
    for v in variants:
        env = dict(os.environ)
        for kv in v.split(","):
            if "=" in kv:
                k, val = kv.split("=", 1)
                env[k] = val
        p = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", workload, str(clips)], env=env, capture_output=True,
                           text=True, timeout=900)
        line = [l for l in p.stdout.splitlines() if l.startswith("AB ")]
        rec = {"workload": workload, "clips": clips, "variant": v}
        rec.update(json.loads(line[0][3:]) if line else {"err": p.stderr[-300:]})
        print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
