"""Cross-process determinism harness: runs the smoke chain (and a tedexp chain) in fresh
processes under different switches and prints a sha256 of every stage, so the switch that changes a hash names the race.

  python profiles/determinism_bisect.py            # parent: spawns the variants, prints one JSON line each
  python profiles/determinism_bisect.py --child beat-ours   # one run, prints {"stage": hash, ...}
"""
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def sha(t):
    return hashlib.sha256(t.detach().float().cpu().contiguous().numpy().tobytes()).hexdigest()[:12]


def child(workload, n_clips, respacing):
    import torch as th
    import gesture_b200  # noqa: F401
    from gesture_b200.engine import chain_for
    from gesture_b200.model_creation import create_model
    from gesture_b200.presets import preset
    from gesture_b200.synthetic import boosted_state_dict, noise_tape, synthetic_wav
    th.cuda.set_device(0)
    params, d_pose, T, L = preset(workload)
    params["Diffusion"]["timestep_respacing"] = respacing
    th.manual_seed(0)
    model, diffusion, *_ = create_model(d_pose, params)
    model.eval()
    model.load_state_dict(boosted_state_dict(model.state_dict(), seed=1))
    N = n_clips
    wav = synthetic_wav(N, L, seed=4)
    x_T, tape = noise_tape((N, d_pose, T), diffusion.num_timesteps, seed=6)
    model.to("cuda:0")
    out = {}
    chain = chain_for(model, diffusion, (N, d_pose, T), "ddpm", "cuda:0")
    chain.begin(x_T.cuda(), wav.cuda(), noise_tape=tape)
    th.cuda.synchronize()
    cond = chain._buffers[5]
    for k, v in cond.items():
        if isinstance(v, th.Tensor):
            out["cond." + k] = sha(v)
    out["tape"] = sha(chain.tape)
    res = chain.run()
    th.cuda.synchronize()
    out["x"] = sha(res["sample"])
    out["eps"] = sha(res["eps"])
    # second chain in the same process (graph replay)
    chain.begin(x_T.cuda(), wav.cuda(), noise_tape=tape)
    res = chain.run()
    th.cuda.synchronize()
    out["x_again"] = sha(res["sample"])
    print("HASH " + json.dumps(out))


VARIANTS = [
    ("default", {}), ("default", {}), ("default", {}),
    ("pdl0", {"GD_PDL": "0"}), ("pdl0", {"GD_PDL": "0"}),
    ("blocking", {"CUDA_LAUNCH_BLOCKING": "1"}),
    ("nograph", {"GD_GRAPH": "0"}), ("nograph_pdl0", {"GD_GRAPH": "0", "GD_PDL": "0"}),
    ("speech_torch", {"GD_SPEECH": "torch"}), ("speech_torch", {"GD_SPEECH": "torch"}),
    ("noconc", {"GD_CONCURRENT": "0"}),
]


def main():
    if "--child" in sys.argv:
        i = sys.argv.index("--child")
        child(sys.argv[i + 1], int(sys.argv[i + 2]), sys.argv[i + 3])
        return
    cases = [("beat-ours", 2, "ddim10"), ("tedexp-ours", 2, "ddim10"), ("beat-ours", 300, "ddim10"), ("tedexp-ours", 64, "ddim10")]
    if len(sys.argv) > 1:
        cases = [c for c in cases if c[0] in sys.argv[1:]]
    for wl, n, rs in cases:
        for name, env in VARIANTS:
            e = dict(os.environ)
            e.update(env)
            p = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", wl, str(n), rs], env=e,
                               capture_output=True, text=True, timeout=900)
            line = [l for l in p.stdout.splitlines() if l.startswith("HASH ")]
            rec = {"workload": wl, "clips": n, "variant": name, "rc": p.returncode}
            if line:
                rec.update(json.loads(line[0][5:]))
            else:
                rec["err"] = p.stderr[-400:]
            print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
