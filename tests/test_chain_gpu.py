"""CUDA sampling path vs the CPU oracle on freshly seeded inputs (small N: the oracle finishes in seconds)."""
import os

import pytest
import torch as th

from util import build, noise_tape, oracle_tables, rel_l2, synthetic_wav

pytestmark = pytest.mark.gpu

# Stated tolerances: bf16 operands AND bf16-rounded GEMM outputs measured at 6.6-7.0e-3 rel-L2 on the
# reference itself, fp32 GEMM outputs at 4.6-4.9e-3 -> bounds below leave 2-3x headroom.
TOL = {"bf16": 2e-2, "fp32act": 1e-2}


@pytest.mark.parametrize("name", ["beat", "tedexp"])
@pytest.mark.parametrize("precision", ["bf16", "fp32act"])
def test_teacher_forced_eps_and_update(name, precision):
    from oracle import ddpm_oracle as orc
    from gesture_b200.engine import chain_for
    N = 3
    model, diffusion, C, T, L, params = build(name, "boost")
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    tabs = oracle_tables(params.Diffusion)
    wav = synthetic_wav(N, L, seed=11)
    x_T, tape = noise_tape((N, C, T), 1000, seed=3)
    feats = orc.speech_features(sd, wav)
    model.to("cuda")
    model.precision = precision
    chain = chain_for(model, diffusion, (N, C, T), "ddpm", "cuda", use_graph=False)
    chain.begin(x_T.cuda(), wav.cuda(), noise_tape=tape.cuda())
    g = th.Generator().manual_seed(5)
    for i in (999, 500, 37, 0):
        x = th.randn(N, C, T, generator=g) * (1.0 + i / 300.0)
        with th.no_grad():
            eps_ref = orc.denoiser(sd, params.type, params.Decoder.heads, x, th.full((N,), i, dtype=th.long), feats)
        chain.set_state(x.cuda(), i)
        chain.step_eager()
        th.cuda.synchronize()
        err = rel_l2(chain.eps, eps_ref)
        assert err < TOL[precision], f"{name}/{precision} t={i}: eps rel-L2 {err:.3e}"
        # the update itself, given the kernel's own eps, must match the oracle's arithmetic to fp32 rounding
        x_ref, x0_ref = orc.ddpm_step(tabs, i, x, chain.eps.cpu(), tape[999 - i])
        assert rel_l2(chain.x, x_ref) < 1e-6 and rel_l2(chain.x0, x0_ref) < 1e-6
        assert int(chain.step.item()) == i - 1


@pytest.mark.parametrize("name", ["beat", "tedexp"])
def test_conditioning_differentials(name):
    """eps must move with the speech and with t the way the oracle's does (a tolerance test alone cannot see a
    dead conditioning path)."""
    from oracle import ddpm_oracle as orc
    N = 2
    model, diffusion, C, T, L, params = build(name, "boost")
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    heads = params.Decoder.heads
    wav1, wav2 = synthetic_wav(N, L, seed=1), synthetic_wav(N, L, seed=2)
    x = th.randn(N, C, T, generator=th.Generator().manual_seed(9))
    t5, t1 = th.full((N,), 500, dtype=th.long), th.full((N,), 100, dtype=th.long)
    with th.no_grad():
        f1, f2 = orc.speech_features(sd, wav1), orc.speech_features(sd, wav2)
        r11 = orc.denoiser(sd, params.type, heads, x, t5, f1)
        r21 = orc.denoiser(sd, params.type, heads, x, t5, f2)
        r12 = orc.denoiser(sd, params.type, heads, x, t1, f1)
    model.to("cuda")
    g11 = model(x.cuda(), t5.cuda(), wav=wav1.cuda()).cpu()
    g21 = model(x.cuda(), t5.cuda(), wav=wav2.cuda()).cpu()
    g12 = model(x.cuda(), t1.cuda(), wav=wav1.cuda()).cpu()
    d_w_ref, d_t_ref = r21 - r11, r12 - r11
    assert d_w_ref.norm() / r11.norm() > 0.02 and d_t_ref.norm() / r11.norm() > 0.02  # the probes are loud
    lim = 0.15 if name == "beat" else 0.5  # a dead path gives ~1.0; tedexp conditioning is only 3% of eps
    assert rel_l2(g21 - g11, d_w_ref) < lim, rel_l2(g21 - g11, d_w_ref)
    assert rel_l2(g12 - g11, d_t_ref) < lim, rel_l2(g12 - g11, d_t_ref)


@pytest.mark.parametrize("name,alg", [("beat", "ddpm"), ("beat", "ddim"), ("tedexp", "ddpm")])
def test_short_respaced_chain_vs_oracle(name, alg):
    """Whole (respaced, 20-step) chain through the public Generator API, graph replay, against the oracle chain."""
    from oracle import ddpm_oracle as orc
    from gesture_b200.generator import Generator
    N = 2
    model, diffusion, C, T, L, params = build(name, "boost", respacing="ddim20")
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    tabs = oracle_tables(params.Diffusion)
    assert len(tabs["betas"]) == 20 and diffusion.num_timesteps == 20
    wav = synthetic_wav(N, L, seed=21)
    x_T, tape = noise_tape((N, C, T), 20, seed=8)
    ref = orc.sample_chain(sd, params.type, params.Decoder.heads, tabs, x_T, wav, tape, alg=alg)
    model.to("cuda")
    gen = Generator(model, diffusion)
    out = gen.generate_sample((N, C, T), wav, noise=x_T, sample_alg=alg, device="cuda", progress=False,
                              noise_tape=tape if alg == "ddpm" else None)
    assert out.shape == (N, T, C)
    err = rel_l2(out.transpose(1, 2), ref)
    assert err < 2e-2, f"{name}/{alg}: final pose rel-L2 {err:.3e}"
    # replaying the captured graph on the same inputs is bit-identical
    out2 = gen.generate_sample((N, C, T), wav, noise=x_T, sample_alg=alg, device="cuda", progress=False,
                               noise_tape=tape if alg == "ddpm" else None)
    assert th.equal(out, out2)


@pytest.mark.parametrize("name", ["beat", "tedexp"])
def test_fused_layernorm_plan_matches_two_kernel_plan(name):
    """The optional plan with every LayerNorm fused into the residual GEMM in front of it (gd_linear_resid_ln, clusters of
    two CTAs for d=512) against the default plan: same 20-step chain, same inputs."""
    from gesture_b200.generator import Generator
    N = 3
    model, diffusion, C, T, L, params = build(name, "boost", respacing="ddim20", device="cuda")
    wav = synthetic_wav(N, L, seed=22)
    x_T, tape = noise_tape((N, C, T), 20, seed=9)
    gen = Generator(model, diffusion)
    outs = []
    for fuse in (False, True):
        model.fuse_layernorm = fuse
        outs.append(gen.generate_sample((N, C, T), wav, noise=x_T, sample_alg="ddpm", device="cuda", progress=False,
                                        noise_tape=tape).clone())
    model.fuse_layernorm = False
    err = rel_l2(outs[1], outs[0])
    assert err < 5e-3, f"{name}: fused vs two-kernel plan rel-L2 {err:.3e}"


@pytest.mark.parametrize("name", ["beat", "tedexp"])
def test_single_clip_chain_vs_oracle(name):
    """BASELINE config 1 shape: one clip (34 / 40 token rows, far less than one 128-row tile) through the public API."""
    from oracle import ddpm_oracle as orc
    from gesture_b200.generator import Generator
    model, diffusion, C, T, L, params = build(name, "boost", respacing="ddim10")
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    tabs = oracle_tables(params.Diffusion)
    wav = synthetic_wav(1, L, seed=23)
    x_T, tape = noise_tape((1, C, T), 10, seed=10)
    ref = orc.sample_chain(sd, params.type, params.Decoder.heads, tabs, x_T, wav, tape, alg="ddpm")
    model.to("cuda")
    out = Generator(model, diffusion).generate_sample((1, C, T), wav, noise=x_T, sample_alg="ddpm", device="cuda", progress=False,
                                                      noise_tape=tape)
    assert out.shape == (1, T, C)
    err = rel_l2(out.transpose(1, 2), ref)
    assert err < 2e-2, f"{name}: single-clip final pose rel-L2 {err:.3e}"


def test_long_form_beat_4x_length():
    """BASELINE config 5: the beat model at 4x the config's sequence length (T=160, wav 128 000 -> 127 memory tokens):
    160-query / 127- and 160-key attention tiles, same weights."""
    from oracle import ddpm_oracle as orc
    from gesture_b200.engine import chain_for
    N, T, L = 2, 160, 128000
    model, diffusion, C, _, _, params = build("beat", "boost")
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    wav = synthetic_wav(N, L, seed=31)
    feats = orc.speech_features(sd, wav)
    assert [f.shape[1] for f in feats] == [125, 124, 126]
    model.to("cuda")
    chain = chain_for(model, diffusion, (N, C, T), "ddpm", "cuda", use_graph=False)
    x_T, _ = noise_tape((N, C, T), 1, seed=2)
    chain.begin(x_T.cuda(), wav.cuda(), need_tape=False)
    assert chain.Tm == 127
    for i in (999, 3):
        x = th.randn(N, C, T, generator=th.Generator().manual_seed(i))
        with th.no_grad():
            ref = orc.denoiser(sd, params.type, params.Decoder.heads, x, th.full((N,), i, dtype=th.long), feats)
        chain.set_state(x.cuda(), i)
        chain.step_eager()
        assert rel_l2(chain.eps, ref) < TOL["bf16"], rel_l2(chain.eps, ref)


def test_batch_invariance():
    """A clip's result must not depend on the batch it is sampled in (what makes clip sharding exact)."""
    from gesture_b200.generator import Generator
    model, diffusion, C, T, L, params = build("beat", "boost", respacing="ddim10", device="cuda")
    gen = Generator(model, diffusion)
    wav = synthetic_wav(5, L, seed=77)
    x_T, tape = noise_tape((5, C, T), 10, seed=78)
    full = gen.generate_sample((5, C, T), wav, noise=x_T, sample_alg="ddpm", device="cuda", progress=False, noise_tape=tape)
    part = gen.generate_sample((2, C, T), wav[3:], noise=x_T[3:], sample_alg="ddpm", device="cuda", progress=False,
                               noise_tape=tape[:, 3:])
    assert th.equal(full[3:], part)


def test_speech_encoder_implementations_agree():
    """The chain's conditioning through the three encoder implementations: native split precision (default), native plain
    bf16, and the fp32 torch module.  With the boosted weights (a random 34-convolution ResNet amplifies rounding noise)
    split precision stays at the bf16 rounding of the final features; plain bf16 does not."""
    from gesture_b200.engine import chain_for
    model, diffusion, C, T, L, _ = build("beat", "boost", device="cuda")
    N = 5
    wav = synthetic_wav(N, L, seed=21).cuda()
    feats = {}
    for impl in ("native", "native-bf16", "torch"):
        model.speech_impl = impl
        chain = chain_for(model, diffusion, (N, C, T), "ddpm", "cuda", use_graph=False)
        assert chain.speech_impl == impl or os.environ.get("GD_SPEECH")
        feats[impl] = [f.clone() for f in chain._speech_features(wav)]
    del model.speech_impl
    for a, b, c in zip(feats["native"], feats["native-bf16"], feats["torch"]):
        assert a.shape == c.shape and tuple(a.shape[::2]) == (N, model.speech_encoder.wav_proj_layer.out_features)
        assert rel_l2(a, c) < 4e-3, rel_l2(a, c)
        assert rel_l2(b, c) < 0.2
        assert rel_l2(a, c) < rel_l2(b, c)


def test_eval_infer_time_ddim():
    """Generator.eval_infer_time_ddim / gpu_warm_up_ddim (generator.py:16-78) on a 10-step respaced process."""
    from gesture_b200.generator import Generator
    model, diffusion, C, T, L, _ = build("beat", "boost", respacing="ddim10", device="cuda")
    N = 3
    kw = {"wav": synthetic_wav(N, L, seed=2).cuda()}
    gen = Generator(model, diffusion)
    for alg in ("ddim", "ddpm"):
        mean_ms, std_ms = gen.eval_infer_time_ddim((N, C, T), kw, sample_alg=alg, repetitions=3, device="cuda")
        assert 0 < mean_ms < 5e3 and std_ms >= 0


def test_p_sample_dict_has_every_reference_key():
    """p_sample_loop(_progressive) hands back the reference's whole dict (gaussian_diffusion.py:276-285,329): sample, mean,
    variance, log_variance, eps, pred_x_start, raw_x_start - with an in-paint blend so raw_x_start != pred_x_start."""
    import numpy as np
    from gesture_b200.diffusion import InpaintBlend
    N = 2
    model, diffusion, C, T, L, params = build("beat", "boost", respacing="ddim10", device="cuda")
    wav = synthetic_wav(N, L, seed=21)
    x_T, tape = noise_tape((N, C, T), 10, seed=22)
    seed = th.randn(N, T, C, generator=th.Generator().manual_seed(23))
    masks = th.ones(N, T, 1)
    masks[:, 10:] = 0
    blend = InpaintBlend(seed.cuda(), masks.cuda(), 0.575, 10, T)
    keys = {"sample", "mean", "variance", "log_variance", "eps", "pred_x_start", "raw_x_start"}
    x_prev = x_T.cuda()
    A = th.from_numpy(diffusion.sqrt_recip_alphas_cumprod).float()
    B = th.from_numpy(diffusion.sqrt_recipm1_alphas_cumprod).float()
    steps = list(diffusion.p_sample_loop_progressive(model, (N, C, T), noise=x_T, model_kwargs={"wav": wav}, denoise_fn=blend,
                                                     device="cuda", noise_tape=tape))
    assert len(steps) == 10
    for k, out in enumerate(steps):
        i = 9 - k
        assert set(out) == keys
        var = np.float32(diffusion.posterior_variance[i])
        logvar = np.float32(diffusion.posterior_log_variance_clipped[i])
        assert th.equal(out["variance"], th.full_like(out["sample"], float(var)))
        assert th.equal(out["log_variance"], th.full_like(out["sample"], float(logvar)))
        raw = A[i] * x_prev - B[i] * out["eps"]                      # _predict_xstart_from_eps (:287-292)
        assert th.equal(out["raw_x_start"], raw)
        assert th.equal(out["pred_x_start"], blend(out["raw_x_start"]))  # denoise_fn (generator.py:271-280)
        c1 = float(np.float32(diffusion.posterior_mean_coef1[i])), float(np.float32(diffusion.posterior_mean_coef2[i]))
        assert th.equal(out["mean"], c1[0] * out["pred_x_start"] + c1[1] * x_prev)
        sig = th.exp(0.5 * th.tensor(float(logvar)))
        expect = out["mean"] + (sig * tape[k].cuda() if i != 0 else 0.0)
        assert th.equal(out["sample"], expect)
        x_prev = out["sample"]
    last = diffusion.p_sample_loop(model, (N, C, T), noise=x_T, model_kwargs={"wav": wav}, denoise_fn=blend, device="cuda",
                                   noise_tape=tape)
    assert set(last) == keys and th.equal(last["sample"], steps[-1]["sample"]) and th.equal(last["mean"], steps[-1]["mean"])


@pytest.mark.parametrize("name,N", [("beat", 3), ("tedexp", 3), ("beat", 130)])
def test_layernorm_prologue_plan_is_bit_identical_to_default_plan(name, N):
    """LayerNorm-prologue plan (gd_linear_ln_bf16, one stand-alone LayerNorm left per step) vs the default two-kernel plan on
    a 20-step chain: the prologue reproduces gd_layernorm bit for bit, so the final poses must be identical."""
    from gesture_b200.generator import Generator
    model, diffusion, C, T, L, params = build(name, "boost", respacing="ddim20", device="cuda")
    wav = synthetic_wav(N, L, seed=31)
    x_T, tape = noise_tape((N, C, T), 20, seed=32)
    outs = []
    for flag in (False, True):
        model.ln_prologue = flag
        outs.append(Generator(model, diffusion).generate_sample((N, C, T), wav, noise=x_T, sample_alg="ddpm", device="cuda",
                                                                progress=False, noise_tape=tape).clone())
    assert th.isfinite(outs[1]).all()
    assert th.equal(outs[0], outs[1]), f"rel-L2 {rel_l2(outs[1], outs[0]):.3e}"


@pytest.mark.parametrize("alg", ["ddpm", "ddim"])
def test_python_denoise_fn_callable_matches_fused_blend(alg):
    """An arbitrary Python `denoise_fn` (gaussian_diffusion.py:256-257) runs on the eager path - denoiser on the kernels, the
    update as torch ops around the call.  With the in-paint closure written as a plain Python function it must reproduce the
    fused InpaintBlend epilogue bit for bit (same eps, same rounding order)."""
    from gesture_b200.diffusion import InpaintBlend
    N = 3
    model, diffusion, C, T, L, params = build("beat", "boost", respacing="ddim10", device="cuda")
    wav = synthetic_wav(N, L, seed=61)
    x_T, tape = noise_tape((N, C, T), 10, seed=62)
    seed = th.randn(N, T, C, generator=th.Generator().manual_seed(63))
    masks = th.ones(N, T, 1)
    masks[:, 8:] = 0
    blend = InpaintBlend(seed.cuda(), masks.cuda(), 0.6, 8, T)
    loop = diffusion.p_sample_loop if alg == "ddpm" else diffusion.ddim_sample_loop
    kw = {"noise_tape": tape} if alg == "ddpm" else {}
    fused = {k: v.clone() for k, v in loop(model, (N, C, T), noise=x_T, model_kwargs={"wav": wav}, denoise_fn=blend, device="cuda", **kw).items()}
    calls = []

    def closure(x0):  # what generator.py:265-280 builds
        calls.append(1)
        return blend(x0)
    eager = loop(model, (N, C, T), noise=x_T, model_kwargs={"wav": wav}, denoise_fn=closure, device="cuda", **kw)
    assert len(calls) == 10
    for k in ("sample", "pred_x_start", "raw_x_start", "eps"):
        assert th.equal(fused[k], eager[k]), k


def test_layer0_memory_qkv_hoist_is_bit_identical():
    """tedexp: LayerNorm + Q|K|V of the layer-0 memory self-attention evaluated once per chain (+ a per-timestep row-0 table)
    instead of in every step - row-wise kernels, so the poses must not change by a bit (N = 3 and a multi-tile N = 40)."""
    from gesture_b200.engine import release_chains
    from gesture_b200.generator import Generator
    for N in (3, 40):
        model, diffusion, C, T, L, params = build("tedexp", "boost", respacing="ddim20", device="cuda")
        wav = synthetic_wav(N, L, seed=81)
        x_T, tape = noise_tape((N, C, T), 20, seed=82)
        outs = []
        for flag in (False, True):
            model.hoist_mem0 = flag
            release_chains(model)
            outs.append(Generator(model, diffusion).generate_sample((N, C, T), wav, noise=x_T, sample_alg="ddpm", device="cuda",
                                                                    progress=False, noise_tape=tape).clone())
        # second chain on the cached plan (buffers refreshed in place) with other speech
        wav2 = synthetic_wav(N, L, seed=83)
        again = Generator(model, diffusion).generate_sample((N, C, T), wav2, noise=x_T, sample_alg="ddpm", device="cuda",
                                                            progress=False, noise_tape=tape).clone()
        model.hoist_mem0 = False
        release_chains(model)
        ref2 = Generator(model, diffusion).generate_sample((N, C, T), wav2, noise=x_T, sample_alg="ddpm", device="cuda",
                                                           progress=False, noise_tape=tape)
        assert th.equal(outs[0], outs[1]) and th.equal(again, ref2)


def test_plans_are_kept_side_by_side(monkeypatch):
    """toggling the in-paint blend (generate_sequence: window 0 without, windows 1.. with) must not rebuild the
    plan and re-capture the graph every time - the two plans live side by side; results stay bit-identical."""
    from gesture_b200 import engine
    from gesture_b200.generator import Generator
    N = 2
    model, diffusion, C, T, L, params = build("beat", "boost", respacing="ddim20", device="cuda")
    gen = Generator(model, diffusion)
    wav = synthetic_wav(N, L, seed=91)
    x_T, tape = noise_tape((N, C, T), 20, seed=92)
    seedp = th.randn(N, T, C, generator=th.Generator().manual_seed(93))
    masks = th.ones(N, T, 1)
    masks[:, 10:] = 0
    builds = []
    real = engine.SamplingChain._build_plan
    monkeypatch.setattr(engine.SamplingChain, "_build_plan", lambda self, cond: (builds.append(1), real(self, cond))[1])
    plain = dict(noise=x_T, sample_alg="ddpm", device="cuda", progress=False, noise_tape=tape)
    blend = dict(plain, inpaint_poses=seedp, inpaint_masks=masks, trans_factor=0.575, pose_seed_len=10)
    a1 = gen.generate_sample((N, C, T), wav, **plain).clone()
    b1 = gen.generate_sample((N, C, T), wav, **blend).clone()
    a2 = gen.generate_sample((N, C, T), wav, **plain).clone()
    b2 = gen.generate_sample((N, C, T), wav, **blend).clone()
    assert len(builds) == 2, builds
    assert th.equal(a1, a2) and th.equal(b1, b2) and not th.equal(a1, b1)
