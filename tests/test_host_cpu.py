"""CPU tests of the host side: the C-ABI library loads and exports every symbol include/gd_b200.h declares (no compute
calls without a GPU), config schemas, the factory's return contract, state_dict compatibility and loud failure without
CUDA (no CPU fallback)."""
import ctypes
import json
import os
import re

import pytest
import torch as th

from util import ROOT, build
from gesture_b200 import _lib
from gesture_b200.generator import Generator
from gesture_b200.json_config import JsonConfig, normalize_model_config
from gesture_b200.model_creation import create_model
from gesture_b200.presets import BEAT_OURS, TEDEXP_OURS, preset


def test_cabi_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "gd_b200.h")).read()
    declared = set(re.findall(r"^\s*(?:int|const char\*|uint64_t|int64_t)\s+(gd_\w+)\s*\(", header, flags=re.M))
    assert {"gd_linear_bf16", "gd_linear_ddpm", "gd_ddpm_update", "gd_layernorm", "gd_dconv_attention"} <= declared
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} not exported by libgd_b200.so"
    assert _lib.load().gd_abi_version() == 4
    assert _lib.load().gd_launch_count() == 0 or _lib.load().gd_launch_count() > 0


def test_cabi_argument_validation_without_gpu():
    lib = _lib.load()
    assert lib.gd_linear_bf16(None, None) == -1 and b"null" in lib.gd_last_error()
    d = _lib.LinearDesc()
    d.A, d.W, d.M, d.N, d.K, d.lda, d.ldw = 16, 16, 8, 64, 100, 104, 104
    assert lib.gd_linear_bf16(ctypes.byref(d), None) == -1 and b"multiple of 64" in lib.gd_last_error()
    assert lib.gd_layernorm(None, 0, None, None, None, 0, 1, 256, 1e-5, None) == -1
    assert lib.gd_dconv_attention(None, None) == -1
    assert lib.gd_step_add(None, 1, None) == -1
    u = _lib.DdpmDesc()
    assert lib.gd_ddpm_update(ctypes.byref(u), None, None) == -1 and b"non-null" in lib.gd_last_error()


def test_json_config_and_schemas(tmp_path):
    path = tmp_path / "my-run.json"
    path.write_text(json.dumps(TEDEXP_OURS))
    cfg = JsonConfig(str(path))
    assert cfg.Meta.name == "my-run" and cfg.Model.Model.args.d_model == 512
    assert isinstance(cfg.Model, JsonConfig) and cfg.to_dict()["Generate"] == {"pose_seed_len": 4}
    with pytest.raises(KeyError):
        cfg.nope
    flat, d_pose, frames = normalize_model_config(cfg)
    assert (flat.type, flat.d_model, flat.Decoder.type, flat.Decoder.heads, flat.Decoder.n_layers) == ("default", 512, "cross_attention", 8, 10)
    assert flat.Diffusion.diffusion_steps == 1000 and flat.Generate.pose_seed_len == 4 and (d_pose, frames) == (126, 34)
    flat2, d2, f2 = normalize_model_config(JsonConfig(BEAT_OURS))
    assert (flat2.type, flat2.d_model, d2, f2) == ("s2g_v2", 256, 123, 40)
    # the Model block alone (what main.py passes, main.py:105-109) works for both schemas
    assert normalize_model_config(JsonConfig(BEAT_OURS).Model)[0].Decoder.n_layers == 4
    assert normalize_model_config(JsonConfig(TEDEXP_OURS).Model)[0].d_model == 512
    merged = JsonConfig({"a": {"x": 1}}) + JsonConfig({"a": {"y": 2}, "b": 3})
    assert merged.a.x == 1 and merged.a.y == 2 and merged.b == 3


@pytest.mark.parametrize("name,n_tensors,n_params", [("beat-ours", 479, 10_336_987), ("tedexp-ours", 919, None)])
def test_create_model_contract(name, n_tensors, n_params):
    params, d_pose, T, L = preset(name)
    model, diffusion, optimizer, sampler, lr_sched = create_model(d_pose, params, lr=1e-3, weight_decay=0.0)
    assert isinstance(model, th.nn.Module) and len(model.state_dict()) == n_tensors  # 479 / 919 tensors
    assert diffusion.num_timesteps == 1000 and diffusion.timestep_map == list(range(1000))
    assert isinstance(optimizer, th.optim.AdamW) and sampler.weights().shape == (1000,)
    t, w = sampler.sample(4, "cpu")
    assert t.shape == (4,) and th.allclose(w, th.ones(4))
    assert "positional_encodings" not in " ".join(model.state_dict())  # non-persistent buffer upstream too
    n = model.count_learnable_parameters()
    assert n > 0 and (n_params is None or abs(n - n_params) < 200_000)
    # a reference-format checkpoint loads strictly and invalidates the packed weights
    v0 = model.weights_version
    model.load_state_dict({k: v.clone() for k, v in model.state_dict().items()}, strict=True)
    assert model.weights_version > v0
    with pytest.raises(NotImplementedError):
        bad = JsonConfig(params.to_dict())
        bad["Decoder"]["type"] = "unet_attention"
        create_model(d_pose, bad)


def test_no_cpu_fallback():
    model, diffusion, C, T, L, params = build("beat", "init")
    gen = Generator(model, diffusion)
    with pytest.raises(_lib.GdError, match="CUDA"):
        gen.generate_sample((1, C, T), th.zeros(1, L), sample_alg="ddpm", device="cpu", progress=False)
    with pytest.raises(_lib.GdError, match="CUDA"):
        model(th.zeros(1, C, T), th.zeros(1, dtype=th.long), wav=th.zeros(1, L))
    with pytest.raises(ValueError):
        gen.generate_sample((1, C, T), th.zeros(1, L), sample_alg="euler", device="cpu", progress=False)
    with pytest.raises(AssertionError):
        gen.generate_sample((1, C, T), th.zeros(L), sample_alg="ddpm", device="cpu", progress=False)
    assert Generator.tensor2dtype(th.ones(2), "array").tolist() == [1.0, 1.0]
    with pytest.raises(ValueError):
        Generator.tensor2dtype(th.ones(2), "list")


def test_inpaint_blend_matches_reference_formula():
    """InpaintBlend as torch op == the reference closure (generator.py:255-281) written out; beat ramp 0.575 -> 1."""
    from gesture_b200.diffusion import InpaintBlend
    N, T, C, seed_len = 2, 40, 5, 10
    g = th.Generator().manual_seed(0)
    seed, x0 = th.randn(N, T, C, generator=g), th.randn(N, C, T, generator=g)
    masks = th.ones(N, T, 1)
    masks[:, seed_len:] = 0
    blend = InpaintBlend(seed, masks, 0.575, seed_len, T)
    assert blend.factor.shape == (T,) and abs(blend.factor[0].item() - 0.575) < 1e-7 and (blend.factor[seed_len:] == 1).all()
    tf = th.cat([th.arange(0.575, 1, (1 - 0.575) / seed_len)[None, :, None], th.ones(1, T - seed_len, 1)], dim=1)
    p = x0.transpose(1, 2)
    ref = ((1 - tf) * masks * seed + tf * masks * p + (1 - masks) * p).transpose(1, 2)
    assert th.equal(blend(x0), ref)
    hard = InpaintBlend(seed, masks, None, None, T)  # trans_factor None -> seed frames copied
    assert th.equal(hard(x0).transpose(1, 2)[:, :seed_len], seed[:, :seed_len])


def test_bench_reference_arm_prints_one_json_line():
    """Driver contract: `bench.py --impl reference` runs on the host alone and prints exactly one JSON line on stdout
    (library banners must not leak into it), with the keys of the reference arm."""
    import json
    import subprocess
    import sys as _sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([_sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["value"] > 0
    from oracle import ref_runner
    # the unmodified reference when oracle/_ref is staged (oracle/make_ref.py), else the oracle port
    assert d["cpu_baseline"]["kind"] == ("reference" if ref_runner.available() else "port")
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["cpu_baseline"]["cores"] >= 1


def test_ddim_eta_tables_reduce_to_the_deterministic_sampler():
    """step_tables('ddim', eta): eta = 0 gives sigma = 0 and the round-1 coefficients; eta = 1 with every step kept makes the
    DDIM variance the ancestral posterior variance (gaussian_diffusion.py:463-467 vs :120-124)."""
    import numpy as np
    from gesture_b200.diffusion import GaussianSpacedDiffusion, get_named_beta_schedule
    d = GaussianSpacedDiffusion(use_timesteps=range(1000), betas=get_named_beta_schedule("linear", 1000), model_var_type="fixed_small")
    A, B, C1, C2, sig = d.step_tables("ddim", 0.0)
    assert float(sig.abs().max()) == 0.0
    sp, sq = np.sqrt(d.alphas_cumprod_prev), np.sqrt(1.0 - d.alphas_cumprod_prev)
    assert th.equal(C1, th.from_numpy(sp - sq / d.sqrt_recipm1_alphas_cumprod).float())
    *_, sig1 = d.step_tables("ddim", 1.0)
    assert th.allclose(sig1[1:] ** 2, th.from_numpy(d.posterior_variance).float()[1:], rtol=1e-4)


def test_inpaint_blend_slices_own_their_memory():
    """InpaintBlend clones its inputs (refreshing a cached plan must not write through to the caller's tensors)
    and `slice` hands every sub-chain an independent copy of its clips."""
    from gesture_b200.diffusion import InpaintBlend
    seed, masks = th.randn(4, 6, 3), th.ones(4, 6, 1)
    b = InpaintBlend(seed, masks, 0.5, 2, 6)
    b.seed.zero_()
    assert seed.abs().sum() > 0
    b = InpaintBlend(seed, masks, 0.5, 2, 6)
    s = b.slice(1, 3)
    assert th.equal(s.seed, seed[1:3]) and s.mask.shape == (2, 6) and th.equal(s.factor, b.factor)
    s.seed.zero_()
    assert th.equal(b.seed, seed)


def test_sub_chain_count_policy(monkeypatch):
    from gesture_b200 import engine

    class M:
        pass
    m = M()
    monkeypatch.delenv("GD_SUBCHAINS", raising=False)
    assert engine.sub_chain_count(m, (128, 123, 40), "ddpm") == 1        # auto: splitting is off (measured: no gain)
    m.sub_chains = 4
    assert engine.sub_chain_count(m, (128, 123, 40), "ddpm") == 4
    assert engine.sub_chain_count(m, (3, 123, 40), "ddpm") == 3          # never more parts than clips
    monkeypatch.setenv("GD_SUBCHAINS", "2")
    assert engine.sub_chain_count(m, (128, 123, 40), "ddpm") == 2
