"""Generates the golden vectors under tests/golden/ by running the REAL reference (imported from /root/reference,
CPU, fp32) on deterministic synthetic weights, speech and noise.  Run in the build container only:

    python tests/golden/make_golden.py [beat|tedexp|tables|all]

Requirements found by the survey: stub `fasttext` (imported transitively, never used on this path), do not import
main.py / datasets, flatten the legacy tedexp config.  Nothing in tests/ or the product reads /root/reference at
test time; only these committed .npz/.pt outputs travel.
"""
import os
import sys
import time
import types

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import gesture_b200  # noqa: E402,F401
from gesture_b200.synthetic import boosted_state_dict, noise_tape, state_dict_digest, synthetic_wav  # noqa: E402

sys.path.insert(0, "/root/reference")
sys.modules["fasttext"] = types.ModuleType("fasttext")
import numpy as np  # noqa: E402
import torch as th  # noqa: E402
from models.generator import Generator  # noqa: E402
from models.model_creation import create_model  # noqa: E402
from utils.json_config import JsonConfig  # noqa: E402

N = 2


def ref_params(name, respacing=""):
    if name == "beat":
        mp = JsonConfig("/root/reference/configs/beat-ours.json").Model
        d_pose, T, L = 123, 40, 32000
    else:
        raw = JsonConfig("/root/reference/configs/tedexp-ours.json")
        M = raw.Model
        mp = JsonConfig({"type": M.Model.type, **M.Model.args, "Encoder": {"type": M.Encoder.type, **M.Encoder.args},
                         "Decoder": {"type": M.Decoder.type, **M.Decoder.args},
                         "Diffusion": {"type": M.Diffusion.type, **M.Diffusion.args}, "Generate": dict(raw.Generate)})
        d_pose, T, L = 126, 34, 36266
    mp["Diffusion"]["timestep_respacing"] = respacing
    return mp, d_pose, T, L


def run_chain(model, diffusion, shape, wav, x_T, tape, alg, denoise_fn=None, keep=()):
    """Unmodified reference loop with th.randn_like replaced by the fixed tape (one draw per step, in order)."""
    it = iter(tape)
    real = th.randn_like
    th.randn_like = lambda x: next(it)
    rec = {}
    loop = diffusion.p_sample_loop_progressive if alg == "ddpm" else diffusion.ddim_sample_loop_progressive
    try:
        n = diffusion.num_timesteps
        x_in = x_T
        for k, out in enumerate(loop(model, shape, noise=x_T, denoise_fn=denoise_fn, model_kwargs={"wav": wav}, device="cpu")):
            i = n - 1 - k
            if i in keep:
                rec[i] = (x_in.clone(), out["eps"].clone(), out["sample"].clone())
            x_in = out["sample"]
    finally:
        th.randn_like = real
    return x_in, rec


def tables():
    out = {}
    for tag, resp in (("full", ""), ("ddim50", "ddim50"), ("sec", "10,15,20")):
        mp, d_pose, _, _ = ref_params("beat", resp)
        _, diff, *_ = create_model(d_pose=d_pose, model_params=mp, is_training=False)
        for k in ("betas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod",
                  "posterior_variance", "posterior_log_variance_clipped", "posterior_mean_coef1", "posterior_mean_coef2"):
            out[f"{tag}.{k}"] = getattr(diff, k)
        out[f"{tag}.timestep_map"] = np.array(diff.timestep_map, dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, "schedule_tables.npz"), **out)
    print("tables written", len(out))


def config(name):
    t0 = time.time()
    out = {}
    for wtag in ("init", "boost"):
        mp, d_pose, T, L = ref_params(name)
        th.manual_seed(0)
        model, diffusion, *_ = create_model(d_pose=d_pose, model_params=mp, is_training=False)
        model.eval()
        if wtag == "boost":
            model.load_state_dict(boosted_state_dict(model.state_dict(), seed=1))
        out[f"{wtag}.digest"] = np.array(state_dict_digest(model.state_dict()))
        shape = (N, d_pose, T)
        wav = synthetic_wav(N, L, seed=123)
        x_T, tape = noise_tape(shape, 1000, seed=99)
        with th.no_grad():
            feats = model.speech_encoder(wavform=wav)
        for nm, f in zip(("low", "mid", "high"), feats):
            out[f"{wtag}.feat_{nm}"] = f.numpy()
        # conditioning sensitivity (how loud wav / t are in eps)
        with th.no_grad():
            t5 = th.full((N,), 500, dtype=th.long)
            e0 = model(x_T, t5, wav=wav)
            e_w = model(x_T, t5, wav=synthetic_wav(N, L, seed=7))
            e_t = model(x_T, th.full((N,), 100, dtype=th.long), wav=wav)
        rel = lambda a, b: ((a - b).norm() / b.norm()).item()  # noqa: E731
        print(f"[{name}/{wtag}] eps rms {e0.pow(2).mean().sqrt():.3f}  d(wav) {rel(e_w, e0):.3e}  d(t) {rel(e_t, e0):.3e}")
        out[f"{wtag}.eps_t500_otherwav"] = e_w.numpy()
        out[f"{wtag}.eps_t100"] = e_t.numpy()
        # full 1000-step ancestral chain, checkpoints along the way (x entering step i, eps, x leaving)
        keep = (999, 998, 500, 20, 1, 0)
        final, rec = run_chain(model, diffusion, shape, wav, x_T, tape, "ddpm", keep=keep)
        for i, (xi, eps, xo) in rec.items():
            out[f"{wtag}.ddpm.x_in.{i}"], out[f"{wtag}.ddpm.eps.{i}"], out[f"{wtag}.ddpm.x_out.{i}"] = xi.numpy(), eps.numpy(), xo.numpy()
        out[f"{wtag}.ddpm.final"] = final.numpy()
        print(f"[{name}/{wtag}] full chain done {time.time() - t0:.0f}s  |final| {final.abs().max():.2f}")
        if wtag == "boost":
            # respaced processes through the public Generator API: ddim50 (DDIM, the reference's default sampler)
            # and ddim50 spacing with the ancestral sampler, plus an in-painted window
            mp2, *_ = ref_params(name, "ddim50")
            th.manual_seed(0)
            m2, d2, *_ = create_model(d_pose=d_pose, model_params=mp2, is_training=False)
            m2.eval()
            m2.load_state_dict(model.state_dict())
            gen = Generator(m2, d2)
            out["boost.ddim50.final"] = gen.generate_sample(shape, wav, noise=x_T, sample_alg="ddim", device="cpu", progress=False).numpy()
            x50, tape50 = noise_tape(shape, 50, seed=5)
            real = th.randn_like
            feed = {"it": iter(tape50)}
            th.randn_like = lambda x: next(feed["it"])
            try:
                out["boost.ddpm50.final"] = gen.generate_sample(shape, wav, noise=x50, sample_alg="ddpm", device="cpu", progress=False).numpy()
                seedp = th.randn(N, T, d_pose, generator=th.Generator().manual_seed(17))
                masks = th.ones(N, T, 1)
                seed_len = mp.Generate.pose_seed_len
                masks[:, seed_len:] = 0
                feed["it"] = iter(tape50)
                out["boost.ddpm50_inpaint.final"] = gen.generate_sample(
                    shape, wav, noise=x50, inpaint_poses=seedp, inpaint_masks=masks, sample_alg="ddpm", trans_factor=0.575,
                    pose_seed_len=seed_len, device="cpu", progress=False).numpy()
                feed["it"] = iter(tape50)  # DDIM draws (and discards) one randn_like per step too
                out["boost.ddim50_inpaint.final"] = gen.generate_sample(
                    shape, wav, noise=x50, inpaint_poses=seedp, inpaint_masks=masks, sample_alg="ddim", trans_factor=None,
                    pose_seed_len=seed_len, device="cpu", progress=False).numpy()
            finally:
                th.randn_like = real
            print(f"[{name}/boost] respaced chains done {time.time() - t0:.0f}s")
    np.savez_compressed(os.path.join(HERE, f"{name}_golden.npz"), **out)
    print(name, "written", sum(v.nbytes for v in out.values()) / 1e6, "MB raw")


def sequence():
    """Long-form windowed generation through the reference's Generator.generate_sequence (generator.py:80-195):
    beat-ours, ddim20 spacing (DDIM sampler, the function's default), 5 s of audio -> 100 frames in 3 windows."""
    mp, d_pose, T, L = ref_params("beat", "ddim20")
    th.manual_seed(0)
    model, diffusion, *_ = create_model(d_pose=d_pose, model_params=mp, is_training=False)
    model.eval()
    model.load_state_dict(boosted_state_dict(model.state_dict(), seed=1))
    gen = Generator(model, diffusion)
    n = 2
    wav_seqs = synthetic_wav(n, 5 * 16000, seed=41)
    g = th.Generator().manual_seed(42)
    init = th.randn(n, 10, d_pose, generator=g)
    x_Ts = [th.randn(n, d_pose, T, generator=g) for _ in range(3)]
    feed = iter(x_Ts)
    real_randn, real_like = th.randn, th.randn_like
    th.randn = lambda *a, **k: next(feed)          # x_T of each window (generate_sample draws it itself)
    th.randn_like = lambda x: th.zeros_like(x)      # DDIM eta=0 multiplies this draw by sigma=0
    try:
        out = {}
        for smooth in (True, False):
            feed = iter(x_Ts)
            res = gen.generate_sequence(wav_seqs, 16000, d_pose, 20, T, 10, smooth_trans=smooth, trans_factor=0.575,
                                        init_poses=init, sample_alg="ddim", device="cpu", progress=False)
            out[f"smooth{int(smooth)}"] = res.numpy()
    finally:
        th.randn, th.randn_like = real_randn, real_like
    np.savez_compressed(os.path.join(HERE, "beat_sequence_golden.npz"), **out)
    print("sequence written", {k: v.shape for k, v in out.items()})


def inpaint():
    """Model type 'inpaint' (Speech2GestureModelInpaint, model.py:120-166; not one of the shipped configs): the tedexp
    decoder with the seed-pose MLP, ddim20 process.  The reference zero-initialises `proj`, which would make the variant
    indistinguishable from 'default'; the boosted weights make it non-trivial.  Saved: proj(...) offset, teacher-forced eps at
    two steps, the in-painted 20-step ancestral chain through Generator.generate_sample."""
    mp, d_pose, T, L = ref_params("tedexp", "ddim20")
    mp["type"] = "inpaint"
    mp["Generate"] = JsonConfig({"pose_seed_len": 4})
    th.manual_seed(0)
    model, diffusion, *_ = create_model(d_pose=d_pose, model_params=mp, is_training=False)
    model.eval()
    model.load_state_dict(boosted_state_dict(model.state_dict(), seed=1))
    n = 2
    wav = synthetic_wav(n, L, seed=61)
    g = th.Generator().manual_seed(62)
    seed_poses = th.randn(n, T, d_pose, generator=g)
    masks = th.ones(n, T, 1)
    masks[:, 4:] = 0
    x_T, tape = noise_tape((n, d_pose, T), diffusion.num_timesteps, seed=63)
    out = {}
    kw = {"wav": wav, "inpaint_pose": seed_poses.transpose(0, 1), "inpaint_mask": masks.transpose(0, 1)}
    with th.no_grad():
        out["offset"] = model.proj(th.cat([kw["inpaint_pose"] * kw["inpaint_mask"], kw["inpaint_mask"]], -1)).permute(1, 2, 0).numpy()
        for i in (19, 3):
            x = th.randn(n, d_pose, T, generator=th.Generator().manual_seed(100 + i))
            t = th.full((n,), diffusion.timestep_map[i], dtype=th.long)
            out[f"eps.{i}"] = model(x, t, **kw).numpy()
    it = iter(tape)
    real = th.randn_like
    th.randn_like = lambda x: next(it)
    try:
        res = Generator(model, diffusion).generate_sample((n, d_pose, T), wav, noise=x_T, inpaint_poses=seed_poses,
                                                          inpaint_masks=masks, sample_alg="ddpm", trans_factor=0.5,
                                                          pose_seed_len=4, device="cpu", progress=False)
    finally:
        th.randn_like = real
    out["ddpm20_inpaint.final"] = res.numpy()
    np.savez_compressed(os.path.join(HERE, "tedexp_inpaint_model_golden.npz"), **out)
    print("inpaint model written", {k: v.shape for k, v in out.items()}, float(np.abs(out["offset"]).mean()))


def bpd():
    """Variational bound through the reference's Generator.eval_bpd (generator.py:197-216 -> calc_bpd_loop): beat-ours,
    20-step respaced process, boosted weights, fixed poses / speech / per-step noise."""
    mp, d_pose, T, L = ref_params("beat", "ddim20")
    th.manual_seed(0)
    model, diffusion, *_ = create_model(d_pose=d_pose, model_params=mp, is_training=False)
    model.eval()
    model.load_state_dict(boosted_state_dict(model.state_dict(), seed=1))
    gen = Generator(model, diffusion)
    n = 2
    wav = synthetic_wav(n, L, seed=51)
    g = th.Generator().manual_seed(52)
    poses = th.randn(n, T, d_pose, generator=g)
    tape = [th.randn(n, d_pose, T, generator=g) for _ in range(diffusion.num_timesteps)]
    it = iter(tape)
    real = th.randn_like
    th.randn_like = lambda x: next(it)
    try:
        res = gen.eval_bpd(poses, wav)
    finally:
        th.randn_like = real
    out = {k: v.numpy() for k, v in res.items()}
    np.savez_compressed(os.path.join(HERE, "beat_bpd_golden.npz"), **out)
    print("bpd written", {k: v.shape for k, v in out.items()}, out["total_bpd"])


def extras():
    """Round-2 boundary cases, beat-ours, boosted weights, through the reference's own entry points:
      * DDIM with eta = 0.5 (ddim_sample_loop, gaussian_diffusion.py:443-484,486-529): 50-step respaced process, fixed tape;
      * one denoiser call with a DIFFERENT timestep per clip (models/model.py:12-15; what training and calc_bpd pass)."""
    mp, d_pose, T, L = ref_params("beat", "ddim50")
    th.manual_seed(0)
    model, diffusion, *_ = create_model(d_pose=d_pose, model_params=mp, is_training=False)
    model.eval()
    model.load_state_dict(boosted_state_dict(model.state_dict(), seed=1))
    n = 3
    shape = (n, d_pose, T)
    wav = synthetic_wav(n, L, seed=71)
    x_T, tape = noise_tape(shape, 50, seed=72)
    out = {}
    it = iter(tape)
    real = th.randn_like
    th.randn_like = lambda x: next(it)
    try:
        with th.no_grad():
            res = diffusion.ddim_sample_loop(model, shape, noise=x_T, model_kwargs={"wav": wav}, device="cpu", eta=0.5)
    finally:
        th.randn_like = real
    out["ddim50_eta05.final"] = res["sample"].numpy()
    t = th.tensor([diffusion.timestep_map[k] for k in (49, 7, 0)], dtype=th.long)  # original timesteps, one per clip
    x = th.randn(shape, generator=th.Generator().manual_seed(73))
    with th.no_grad():
        out["eps_per_clip_t"] = model(x, t, wav=wav).numpy()
    out["per_clip_t"] = t.numpy()
    np.savez_compressed(os.path.join(HERE, "beat_extras_golden.npz"), **out)
    print("extras written", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    th.set_num_threads(8)
    if what in ("tables", "all"):
        tables()
    if what in ("beat", "all"):
        config("beat")
    if what in ("tedexp", "all"):
        config("tedexp")
    if what in ("sequence", "all"):
        sequence()
    if what in ("bpd", "all"):
        bpd()
    if what in ("inpaint", "all"):
        inpaint()
    if what in ("extras", "all"):
        extras()
