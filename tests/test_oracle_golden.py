"""CPU tests (no GPU): the oracle against the golden vectors produced by the REAL reference
(tests/golden/make_golden.py), and the package's host logic (tables, config schemas, weights) against both."""
import hashlib

import numpy as np
import pytest
import torch as th

from util import GOLDEN, build, load_golden, noise_tape, rel_l2, state_dict_digest, synthetic_wav
from oracle import ddpm_oracle as orc

TABLE_KEYS = ("betas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod",
              "posterior_variance", "posterior_log_variance_clipped", "posterior_mean_coef1", "posterior_mean_coef2")


@pytest.fixture(scope="module")
def tables():
    return np.load(f"{GOLDEN}/schedule_tables.npz")


@pytest.mark.parametrize("tag,resp", [("full", ""), ("ddim50", "ddim50"), ("sec", "10,15,20")])
def test_schedule_tables_bit_exact(tables, tag, resp):
    """float64 tables: oracle == package == reference, bit for bit (betas are re-derived)."""
    from gesture_b200.diffusion import create_diffusion
    from gesture_b200.json_config import JsonConfig
    o = orc.spaced_diffusion_tables("linear", 1000, resp)
    d = create_diffusion(JsonConfig({"type": "gaussian", "noise_schedule": "linear", "diffusion_steps": 1000,
                                     "timestep_respacing": resp, "model_var_type": "fixed_small"}), False)
    for k in TABLE_KEYS:
        g = tables[f"{tag}.{k}"]
        assert g.dtype == np.float64
        assert np.array_equal(o[k], g), f"oracle {tag}.{k}"
        assert np.array_equal(getattr(d, k), g), f"package {tag}.{k}"
    assert np.array_equal(o["timestep_map"], tables[f"{tag}.timestep_map"])
    assert d.timestep_map == list(tables[f"{tag}.timestep_map"])


def test_schedule_known_answers(tables):
    """Known answers recorded by the survey from the live reference."""
    t = orc.spaced_diffusion_tables("linear", 1000, "")
    assert t["betas"][0] == 9.9999999999988987e-05 and t["betas"][0] != 1e-4
    assert t["betas"][999] == 0.020000000000000018
    assert t["alphas_cumprod"][999] == 4.0358297653756761e-05
    assert t["sqrt_recip_alphas_cumprod"][999] == 157.41045725150062
    assert t["sqrt_recipm1_alphas_cumprod"][0] == 0.01000050003749898
    assert t["posterior_variance"][0] == 0 and t["posterior_variance"][999] == 0.019999983526560627
    assert t["posterior_log_variance_clipped"][0] == -9.8167251352959504
    assert t["posterior_mean_coef1"][0] == 1 and t["posterior_mean_coef2"][0] == 0
    assert t["posterior_mean_coef2"][999] == 0.98994867826751731
    for key, h64 in (("betas", "064e84ed82ea1a4c"), ("alphas_cumprod", "781933865cefafac"),
                     ("posterior_mean_coef2", "2693319332724041")):
        assert hashlib.sha256(t[key].tobytes()).hexdigest()[:16] == h64


def test_step_tables_fp32():
    """The five fp32 per-step coefficient tables handed to the kernels are the reference's float64 tables cast with
    .float() (gaussian_diffusion.py:691); sigma = exp(0.5*logvar) evaluated in fp32 like p_sample does."""
    from gesture_b200.presets import preset
    from gesture_b200.diffusion import create_diffusion
    params, *_ = preset("beat-ours")
    d = create_diffusion(params.Diffusion, False)
    A, B, C1, C2, sg = d.step_tables("ddpm")
    t = orc.spaced_diffusion_tables("linear", 1000, "")
    assert th.equal(A, th.from_numpy(t["sqrt_recip_alphas_cumprod"]).float())
    assert th.equal(C2, th.from_numpy(t["posterior_mean_coef2"]).float())
    assert th.equal(sg, th.exp(0.5 * th.from_numpy(t["posterior_log_variance_clipped"]).float()))
    # DDIM as the same affine update: x0*sqrt(ab_prev) + sqrt(1-ab_prev)*(A x - x0)/B
    A, B, C1, C2, sg = d.step_tables("ddim")
    x, eps = th.randn(4, 7, dtype=th.float64), th.randn(4, 7, dtype=th.float64)
    for i in (999, 500, 1, 0):
        a, b = t["sqrt_recip_alphas_cumprod"][i], t["sqrt_recipm1_alphas_cumprod"][i]
        x0 = a * x - b * eps
        ref = x0 * np.sqrt(t["alphas_cumprod_prev"][i]) + np.sqrt(1 - t["alphas_cumprod_prev"][i]) * (a * x - x0) / b
        got = C1[i].double() * x0 + C2[i].double() * x
        assert (got - ref).abs().max() < 1e-5 * max(1.0, ref.abs().max())
    assert (sg == 0).all()


@pytest.mark.parametrize("name", ["beat", "tedexp"])
@pytest.mark.parametrize("weights", ["init", "boost"])
def test_weights_reproduce_reference(name, weights):
    """Same seed -> same random init as the reference's create_model (digest recorded from the real reference)."""
    g = load_golden(name)
    model, *_ = build(name, weights)
    assert state_dict_digest(model.state_dict()) == str(g[f"{weights}.digest"])


@pytest.mark.parametrize("name", ["beat", "tedexp"])
def test_oracle_speech_features_and_eps(name):
    g = load_golden(name)
    for weights in ("init", "boost"):
        model, diffusion, C, T, L, params = build(name, weights)
        sd = model.state_dict()
        wav = synthetic_wav(2, L, seed=123)
        with th.no_grad():
            feats = orc.speech_features(sd, wav)
            mine = model.speech_encoder(wavform=wav)
        for nm, f, m in zip(("low", "mid", "high"), feats, mine):
            assert rel_l2(f, g[f"{weights}.feat_{nm}"]) < 1e-5, (weights, nm)
            assert rel_l2(m, g[f"{weights}.feat_{nm}"]) < 1e-5, (weights, nm)  # package's torch speech encoder
        heads = params.Decoder.heads
        for i in (999, 500, 0):
            x = th.from_numpy(g[f"{weights}.ddpm.x_in.{i}"])
            with th.no_grad():
                eps = orc.denoiser(sd, params.type, heads, x, th.full((2,), i, dtype=th.long), feats)
            assert rel_l2(eps, g[f"{weights}.ddpm.eps.{i}"]) < 1e-3, (weights, i)
        if weights == "boost":
            x_T, _ = noise_tape((2, C, T), 1000, seed=99)
            with th.no_grad():
                e_t = orc.denoiser(sd, params.type, heads, x_T, th.full((2,), 100, dtype=th.long), feats)
                e_w = orc.denoiser(sd, params.type, heads, x_T, th.full((2,), 500, dtype=th.long),
                                   orc.speech_features(sd, synthetic_wav(2, L, seed=7)))
            assert rel_l2(e_t, g["boost.eps_t100"]) < 1e-3 and rel_l2(e_w, g["boost.eps_t500_otherwav"]) < 1e-3


@pytest.mark.parametrize("name", ["beat", "tedexp"])
def test_oracle_chain_tail_and_update(name):
    """Last 21 steps of the reference's 1000-step chain replayed by the oracle from the recorded x entering i=20;
    single updates at i=999/998 must match the reference's arithmetic to fp32 rounding."""
    g = load_golden(name)
    model, diffusion, C, T, L, params = build(name, "boost")
    sd = model.state_dict()
    tabs = orc.spaced_diffusion_tables("linear", 1000, "")
    _, tape = noise_tape((2, C, T), 1000, seed=99)
    for i in (999, 998, 500, 1, 0):
        x_next, _ = orc.ddpm_step(tabs, i, th.from_numpy(g[f"boost.ddpm.x_in.{i}"]), th.from_numpy(g[f"boost.ddpm.eps.{i}"]), tape[999 - i])
        assert th.equal(x_next, th.from_numpy(g[f"boost.ddpm.x_out.{i}"])), i
    if name == "tedexp":
        return  # the 21-step tail costs ~10 s for beat, ~1 min for tedexp: keep the CPU suite short
    wav = synthetic_wav(2, L, seed=123)
    x = th.from_numpy(g["boost.ddpm.x_in.20"])
    feats = orc.speech_features(sd, wav)
    with th.no_grad():
        for i in range(20, -1, -1):
            eps = orc.denoiser(sd, params.type, params.Decoder.heads, x, th.full((2,), i, dtype=th.long), feats)
            x, _ = orc.ddpm_step(tabs, i, x, eps, tape[999 - i])
    assert rel_l2(x, g["boost.ddpm.final"]) < 1e-3


def test_oracle_respaced_and_inpaint_chains():
    """ddim50 process: DDIM sampler, ancestral sampler and both with the in-paint blend, vs the reference's
    Generator.generate_sample outputs (beat; 50 steps each)."""
    g = load_golden("beat")
    model, diffusion, C, T, L, params = build("beat", "boost", respacing="ddim50")
    sd = model.state_dict()
    tabs = orc.spaced_diffusion_tables("linear", 1000, "ddim50")
    wav = synthetic_wav(2, L, seed=123)
    x_T, _ = noise_tape((2, C, T), 1000, seed=99)
    heads = params.Decoder.heads
    out = orc.sample_chain(sd, params.type, heads, tabs, x_T, wav, None, alg="ddim")
    assert rel_l2(out.transpose(1, 2), g["boost.ddim50.final"]) < 1e-3
    x50, tape50 = noise_tape((2, C, T), 50, seed=5)
    out = orc.sample_chain(sd, params.type, heads, tabs, x50, wav, tape50, alg="ddpm")
    assert rel_l2(out.transpose(1, 2), g["boost.ddpm50.final"]) < 1e-3
    seedp = th.randn(2, T, C, generator=th.Generator().manual_seed(17))
    masks = th.ones(2, T, 1)
    masks[:, 10:] = 0
    f = orc.transition_factor(0.575, 10, T)
    out = orc.sample_chain(sd, params.type, heads, tabs, x50, wav, tape50, alg="ddpm",
                           blend=lambda x0: orc.inpaint_blend(x0, seedp, masks, f))
    assert rel_l2(out.transpose(1, 2), g["boost.ddpm50_inpaint.final"]) < 1e-3
    out = orc.sample_chain(sd, params.type, heads, tabs, x50, wav, None, alg="ddim",
                           blend=lambda x0: orc.inpaint_blend(x0, seedp, masks, 0))
    assert rel_l2(out.transpose(1, 2), g["boost.ddim50_inpaint.final"]) < 1e-3


def test_oracle_bpd_terms():
    """Variational-bound terms (calc_bpd_loop) of the oracle vs the reference's Generator.eval_bpd: beat, ddim20 process."""
    import numpy as np
    from util import GOLDEN
    g = np.load(f"{GOLDEN}/beat_bpd_golden.npz")
    model, diffusion, C, T, L, params = build("beat", "boost", respacing="ddim20")
    tabs = orc.spaced_diffusion_tables("linear", 1000, "ddim20")
    wav = synthetic_wav(2, L, seed=51)
    rg = th.Generator().manual_seed(52)
    poses = th.randn(2, T, C, generator=rg)
    tape = [th.randn(2, C, T, generator=rg) for _ in range(20)]
    out = orc.bpd_loop(model.state_dict(), params.type, params.Decoder.heads, tabs, poses.transpose(1, 2), wav, tape)
    for k in ("total_bpd", "prior_bpd", "x_start_mse", "vb", "mse"):
        assert rel_l2(out[k], g[k]) < 1e-3, k


def test_oracle_inpaint_model_variant():
    """Model type 'inpaint' (seed-pose MLP offset on the denoiser input): oracle vs the reference's own model and its
    Generator.generate_sample (tedexp decoder, ddim20 process, boosted weights incl. a non-zero `proj`)."""
    import numpy as np
    from util import GOLDEN
    g = np.load(f"{GOLDEN}/tedexp_inpaint_model_golden.npz")
    model, diffusion, C, T, L, params = build("tedexp", "boost", respacing="ddim20", model_type="inpaint")
    sd = model.state_dict()
    assert {"proj.0.weight", "proj.2.bias", "proj.4.weight"} <= set(sd) and model.pose_seed_len == 4
    tabs = orc.spaced_diffusion_tables("linear", 1000, "ddim20")
    wav = synthetic_wav(2, L, seed=61)
    seed_poses = th.randn(2, T, C, generator=th.Generator().manual_seed(62))
    masks = th.ones(2, T, 1)
    masks[:, 4:] = 0
    off = orc.inpaint_offset(sd, seed_poses, masks)
    assert rel_l2(off, g["offset"]) < 1e-5
    feats = orc.speech_features(sd, wav)
    heads = params.Decoder.heads
    for i in (19, 3):
        x = th.randn(2, C, T, generator=th.Generator().manual_seed(100 + i))
        t = th.full((2,), int(tabs["timestep_map"][i]), dtype=th.long)
        with th.no_grad():
            eps = orc.denoiser(sd, "inpaint", heads, x, t, feats, offset=off)
        assert rel_l2(eps, g[f"eps.{i}"]) < 1e-3
    x_T, tape = noise_tape((2, C, T), 20, seed=63)
    f = orc.transition_factor(0.5, 4, T)
    out = orc.sample_chain(sd, "inpaint", heads, tabs, x_T, wav, tape, alg="ddpm", offset=off,
                           blend=lambda x0: orc.inpaint_blend(x0, seed_poses, masks, f))
    assert rel_l2(out.transpose(1, 2), g["ddpm20_inpaint.final"]) < 1e-3


def test_oracle_ddim_eta_and_per_clip_t_vs_reference():
    """Round-2 boundary cases against tests/golden/beat_extras_golden.npz (written by the unmodified reference): DDIM with
    eta = 0.5 over a 50-step process, and one denoiser call with a different timestep per clip."""
    import numpy as np
    from oracle import ddpm_oracle as orc
    from util import GOLDEN, build, noise_tape, rel_l2, synthetic_wav
    g = np.load(f"{GOLDEN}/beat_extras_golden.npz")
    model, diffusion, C, T, L, params = build("beat", "boost", respacing="ddim50")
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    tabs = orc.spaced_diffusion_tables("linear", 1000, "ddim50")
    wav = synthetic_wav(3, L, seed=71)
    x_T, tape = noise_tape((3, C, T), 50, seed=72)
    out = orc.sample_chain(sd, params.type, params.Decoder.heads, tabs, x_T, wav, tape, alg="ddim", eta=0.5)
    assert rel_l2(out, g["ddim50_eta05.final"]) < 1e-3
    x = th.randn(3, C, T, generator=th.Generator().manual_seed(73))
    with th.no_grad():
        eps = orc.denoiser(sd, params.type, params.Decoder.heads, x, th.from_numpy(g["per_clip_t"]), orc.speech_features(sd, wav))
    assert rel_l2(eps, g["eps_per_clip_t"]) < 1e-4
