"""The two shipped model configurations, restated as dicts (the reference's JSON files also carry dataset paths
and mean-pose arrays that the sampling path never reads).

BEAT_OURS is in the flat schema of `configs/beat-ours.json:59-83`; TEDEXP_OURS keeps the legacy nested schema of
`configs/tedexp-ours.json:6-67` on purpose, so that the schema adapter is exercised by everything that uses it.
"""
from .json_config import JsonConfig, normalize_model_config

_DIFFUSION = {"type": "gaussian", "noise_schedule": "linear", "diffusion_steps": 1000, "timestep_respacing": "",
              "model_var_type": "fixed_small"}

BEAT_OURS = {
    "Data": {"pose_fps": 20, "wav_sr": 16000, "pose_window_len": 40, "pose_stride_len": 20,
             "pose_representation": "log_rot", "joints": [f"joint{i}" for i in range(41)]},
    "Model": {"type": "s2g_v2", "d_model": 256, "dropout_prob": 0.0, "Encoder": {"type": "ha2g"},
              "Decoder": {"type": "oneway_cross_attention", "heads": 8, "n_layers": 4},
              "Diffusion": dict(_DIFFUSION),
              "Generate": {"pose_seed_len": 10, "smooth_transition": False, "trans_factor": 0.575}},
}

TEDEXP_OURS = {
    "Data": {"type": "ted_exp", "args": {"n_poses": 34, "subdivision_stride": 10, "pose_resampling_fps": 15,
                                         "pose_dim": 126}},
    "Model": {"Model": {"type": "default", "args": {"d_model": 512, "dropout_prob": 0.0, "pose_seed_len": 4}},
              "Encoder": {"type": "ha2g", "args": {}},
              "Decoder": {"type": "cross_attention", "args": {"heads": 8, "n_layers": 10}},
              "Diffusion": {"type": "gaussian", "args": {k: v for k, v in _DIFFUSION.items() if k != "type"}}},
    "Generate": {"pose_seed_len": 4},
}

# wav samples per window: beat 40 frames @20 fps of 16 kHz audio; tedexp 34 frames @15 fps (HA2G convention)
WAV_LEN = {"beat-ours": 32000, "tedexp-ours": 36266}


def preset(name):
    """-> (flat model_params JsonConfig, d_pose, n_frames, wav_len) for 'beat-ours' | 'tedexp-ours'."""
    raw = {"beat-ours": BEAT_OURS, "tedexp-ours": TEDEXP_OURS}[name]
    params, d_pose, n_frames = normalize_model_config(JsonConfig(raw))
    return params, d_pose, n_frames, WAV_LEN[name]
