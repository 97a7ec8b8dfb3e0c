"""TEST INFRASTRUCTURE - runs the UNMODIFIED reference sampler staged under oracle/_ref/ (see oracle/make_ref.py) on CPU.

Used by bench.py's CPU arms (`cpu_baseline.kind = "reference"`, `--impl reference`) and by tests; never by the product.
Follows BASELINE.md section 4: `th.manual_seed(0)` -> `create_model(d_pose, params, is_training=False)` -> `model.eval()`
-> `Generator(model, diffusion)`; the tedexp config is flattened to the schema `create_model` reads;
`main.py` / `datasets/` are never imported.
"""
import hashlib
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")

SHAPES = {"beat-ours": (123, 40, 32000), "tedexp-ours": (126, 34, 36266)}


def available():
    return os.path.exists(os.path.join(REF, "MANIFEST.json"))


def verify():
    """Every staged file still has the digest recorded when it was copied from the reference."""
    man = json.load(open(os.path.join(REF, "MANIFEST.json")))
    for rel, digest in man["files"].items():
        got = hashlib.sha256(open(os.path.join(REF, rel), "rb").read()).hexdigest()
        if got != digest:
            raise RuntimeError(f"oracle/_ref/{rel} differs from the reference file it was copied from")
    return len(man["files"])


def _import_reference():
    if not available():
        raise RuntimeError("oracle/_ref is not staged: run `python oracle/make_ref.py` in the build container")
    verify()
    sys.dont_write_bytecode = True
    if REF not in sys.path:
        sys.path.insert(0, REF)  # provides `models`, `utils` and the `fasttext` stub
    from models.generator import Generator
    from models.model_creation import create_model
    from utils.json_config import JsonConfig
    return Generator, create_model, JsonConfig


def build(workload, respacing=""):
    """-> (generator, model, diffusion, d_pose, T, wav_len) of the reference for 'beat-ours' | 'tedexp-ours'."""
    import torch as th
    Generator, create_model, JsonConfig = _import_reference()
    d_pose, T, L = SHAPES[workload]
    if workload == "beat-ours":
        mp = JsonConfig(os.path.join(REF, "configs", "beat-ours.json")).Model
    else:
        raw = JsonConfig(os.path.join(REF, "configs", "tedexp-ours.json"))
        M = raw.Model
        mp = JsonConfig({"type": M.Model.type, **M.Model.args, "Encoder": {"type": M.Encoder.type, **M.Encoder.args},
                         "Decoder": {"type": M.Decoder.type, **M.Decoder.args},
                         "Diffusion": {"type": M.Diffusion.type, **M.Diffusion.args}, "Generate": dict(raw.Generate)})
    mp["Diffusion"]["timestep_respacing"] = respacing
    th.manual_seed(0)
    model, diffusion, *_ = create_model(d_pose=d_pose, model_params=mp, is_training=False)
    model.eval()
    return Generator(model, diffusion), model, diffusion, d_pose, T, L


def time_chain(workload, clips, denoise_steps, threads=None, wav_seed=123, noise_seed=99):
    """Time the first `denoise_steps` steps of the reference's own ancestral loop (`p_sample_loop_progressive`, the generator
    behind `Generator.generate_sample(sample_alg='ddpm')`, gaussian_diffusion.py:365-412) exactly as shipped - the speech
    encoder is re-run inside every denoiser call (models/model.py:54-56) - and extrapolate to the 1000-step chain."""
    import torch as th
    th.set_num_threads(threads or os.cpu_count() or 1)
    gen, model, diffusion, C, T, L = build(workload)
    wav = th.randn(clips, L, generator=th.Generator().manual_seed(wav_seed))
    x_T = th.randn(clips, C, T, generator=th.Generator().manual_seed(noise_seed))
    n = diffusion.num_timesteps
    steps = min(denoise_steps, n)
    with th.no_grad():
        it = diffusion.p_sample_loop_progressive(model, (clips, C, T), noise=x_T, model_kwargs={"wav": wav}, device="cpu")
        next(it)  # first step untimed: lazy initialisation, allocator warm-up
        t0 = time.perf_counter()
        for _ in range(steps):
            out = next(it)
        dt = time.perf_counter() - t0
    assert bool(th.isfinite(out["sample"]).all())
    per_step = dt / steps
    return {"frames_per_s": clips * T / (per_step * n), "ms_per_denoise_step": per_step * 1e3, "seconds": dt,
            "cores": th.get_num_threads(), "clips": clips, "denoise_steps": steps, "chain_steps": n}
