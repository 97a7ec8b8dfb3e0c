"""Factory with the reference's signature and return tuple (`models/model_creation.py:51-191`)."""
import torch as th

from .diffusion import create_diffusion
from .json_config import JsonConfig, normalize_model_config
from .model import Speech2GestureDenoiser
from .modules import PoseDecoderParams, SpeechEncoder, StepEncoderParams

_UNSUPPORTED_DECODERS = ("cross_attention_gcn", "unet_attention")


class UniformSampler:
    """`create_named_schedule_sampler("uniform", diffusion)` (models/modules/resample.py:12-62): training only."""

    def __init__(self, diffusion):
        self.diffusion = diffusion
        self._weights = th.ones(diffusion.num_timesteps)

    def weights(self):
        return self._weights

    def sample(self, batch_size, device):
        w = self.weights().numpy()
        p = w / w.sum()
        import numpy as np
        idx = np.random.choice(len(p), size=(batch_size,), p=p)
        return th.from_numpy(idx).long().to(device), th.from_numpy(1 / (len(p) * p[idx])).float().to(device)


class ConstantLR(th.optim.lr_scheduler.LambdaLR):
    """models/lr_scheduler.py ConstantLR."""

    def __init__(self, optimizer):
        super().__init__(optimizer, lambda step: 1.0)


def create_lr_scheduler(params, optimizer):
    """models/model_creation.py:19-27. Only the schedules the sampling path can meet are built natively."""
    if params.type == "const":
        return ConstantLR(optimizer)
    if params.type in ("noam", "noamxf"):
        raise NotImplementedError("Noam schedules belong to the training loop, which is out of scope here")
    raise ValueError("Unsupport lr_scheduler type.")


def create_model(d_pose, model_params, lr=1e-2, weight_decay=None, scheduler_params=None, is_training=False):
    """-> (model, diffusion, optimizer, schedule_sampler, lr_scheduler), as the reference.

    `model_params` may be the reference's flat `config.Model` block (beat-ours), the legacy nested block of
    tedexp-ours, or a whole-file config; the nested forms are flattened first.
    """
    if weight_decay is None:
        weight_decay = 0.0
    model_params, _, _ = normalize_model_config(model_params)

    encoder_params = model_params.get("Encoder")
    if encoder_params.type != "ha2g":
        raise ValueError
    speech_encoder = SpeechEncoder(d_model=model_params.d_model, dropout_prob=model_params.dropout_prob)

    decoder_params = model_params.get("Decoder")
    if decoder_params.type in _UNSUPPORTED_DECODERS:
        raise NotImplementedError(f"decoder type {decoder_params.type} is not used by either shipped config and has no "
                                  "B200 kernels")
    if decoder_params.type not in ("cross_attention", "oneway_cross_attention"):
        raise ValueError(f"Unsupported decoder type {decoder_params.type}.")
    decoder = PoseDecoderParams(decoder_params.type, d_pose, model_params.d_model, decoder_params.heads,
                                decoder_params.n_layers, d_pose)
    step_encoder = StepEncoderParams(model_params.d_model, model_params.dropout_prob)
    diffusion = create_diffusion(model_params.get("Diffusion"), is_training)

    if model_params.type not in ("default", "s2g_v2", "inpaint"):
        raise ValueError(f"Unsupported model_type {model_params.type}")
    if model_params.type == "inpaint" and decoder_params.type != "cross_attention":
        raise NotImplementedError("the inpaint wrapper is built on the default (tedexp) memory layout: cross_attention decoder only")
    extra = {}
    if model_params.type == "inpaint":  # model_creation.py:134-143
        extra = {"dropout_prob": model_params.dropout_prob, "pose_seed_len": model_params.Generate.pose_seed_len}
    model = Speech2GestureDenoiser(model_params.type, d_pose, model_params.d_model, speech_encoder, decoder, step_encoder,
                                   **extra)
    model.bind_diffusion(diffusion)

    optimizer = th.optim.AdamW(model.parameters(), lr=lr, weight_decay=weight_decay)
    schedule_sampler = UniformSampler(diffusion)
    if scheduler_params is None:
        scheduler_params = JsonConfig({"type": "const"})
    lr_scheduler = create_lr_scheduler(scheduler_params, optimizer)
    return model, diffusion, optimizer, schedule_sampler, lr_scheduler
