"""Shared helpers for the parity tests."""
import os
import sys

import numpy as np
import torch as th

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")

import gesture_b200  # noqa: E402,F401
from gesture_b200.model_creation import create_model  # noqa: E402
from gesture_b200.presets import preset  # noqa: E402
from gesture_b200.synthetic import boosted_state_dict, noise_tape, state_dict_digest, synthetic_wav  # noqa: E402

SHORT = {"beat": "beat-ours", "tedexp": "tedexp-ours"}


def build(name, weights="boost", respacing="", device="cpu", model_type=None):
    """-> (model, diffusion, d_pose, T, wav_len, flat params).  weights: 'init' (seed-0 default init) | 'boost'.
    model_type='inpaint' switches the wrapper (Speech2GestureModelInpaint, tedexp decoder only)."""
    params, d_pose, T, L = preset(SHORT[name])
    params["Diffusion"]["timestep_respacing"] = respacing
    if model_type is not None:
        params["type"] = model_type
    th.manual_seed(0)
    model, diffusion, *_ = create_model(d_pose, params)
    model.eval()
    if weights == "boost":
        model.load_state_dict(boosted_state_dict(model.state_dict(), seed=1))
    if device != "cpu":
        model.to(device)
    return model, diffusion, d_pose, T, L, params


def oracle_tables(diffusion_params):
    from oracle import ddpm_oracle as orc
    return orc.spaced_diffusion_tables(diffusion_params.noise_schedule, diffusion_params.diffusion_steps,
                                       diffusion_params.timestep_respacing)


def rel_l2(a, b):
    a, b = th.as_tensor(a).float().cpu(), th.as_tensor(b).float().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def load_golden(name):
    path = os.path.join(GOLDEN, f"{name}_golden.npz")
    return np.load(path)
