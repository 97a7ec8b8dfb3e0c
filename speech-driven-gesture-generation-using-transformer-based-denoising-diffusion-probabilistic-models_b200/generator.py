"""Sampler façade with the reference's interface (`models/generator.py:8-309`)."""
from typing import Tuple

import numpy as np
import torch as th

from .diffusion import GaussianSpacedDiffusion, InpaintBlend
from .model import Speech2GestureDenoiser


class Generator:
    def __init__(self, model: Speech2GestureDenoiser, diffusion: GaussianSpacedDiffusion) -> None:
        self.model = model
        self.diffusion = diffusion
        model.bind_diffusion(diffusion)

    def _choose_sample_func(self, sample_alg: str):
        if sample_alg == "ddim":
            return self.diffusion.ddim_sample_loop
        if sample_alg == "ddpm":
            return self.diffusion.p_sample_loop
        raise ValueError(f"Unsupported sample algorithm: {sample_alg}")

    @th.no_grad()
    def generate_sample(
        self,
        shape: Tuple[int],  # (N,C,T)
        wavs: th.Tensor,  # (N,T_wav)
        noise: th.Tensor = None,
        inpaint_poses: th.Tensor = None,  # (N,T,C)
        inpaint_masks: th.Tensor = None,  # (N,T,1)
        sample_alg: str = "ddim",
        trans_factor: float = None,
        pose_seed_len: int = None,
        return_dtype: str = "tensor",  # 'tensor' | 'cpu_tensor' | 'array'
        device: str = "cpu",
        progress: bool = True,
        noise_tape: th.Tensor = None,  # extension: (n_steps, N, C, T) fixed per-step noise, loop order
    ):
        """generator.py:218-296.  The whole reverse chain runs as CUDA-graph replays of the fused step."""
        wavs = wavs.to(device)
        assert len(wavs.shape) == 2, f"Wav dim should be (N,T). Got: {wavs.shape}"
        assert len(shape) == 3, f"Shape should be (N,C,T). Got: {shape}"
        sample_func = self._choose_sample_func(sample_alg)
        denoise_fn = None
        if inpaint_poses is not None:
            assert inpaint_masks is not None, "Provide inpaint_masks for inpainting."
            denoise_fn = InpaintBlend(inpaint_poses.to(device), inpaint_masks.to(device), trans_factor, pose_seed_len,
                                      shape[2])
        if noise is None:
            noise = th.randn(shape, device=device)
        kw = {"noise_tape": noise_tape} if sample_alg == "ddpm" else {}
        model_kwargs = {"wav": wavs}
        if getattr(self.model, "model_type", None) == "inpaint":  # generator.py:244-250
            assert len(inpaint_poses.shape) == 3
            assert len(inpaint_masks.shape) == 3
            assert inpaint_masks.size()[:2] == inpaint_poses.size()[:2]
            model_kwargs["inpaint_pose"] = inpaint_poses.to(device).transpose(0, 1)  # -> (T,N,C)
            model_kwargs["inpaint_mask"] = inpaint_masks.to(device).transpose(0, 1)  # -> (T,N,1)
        sample = sample_func(self.model, shape, noise=noise, denoise_fn=denoise_fn, model_kwargs=model_kwargs,
                             device=device, progress=progress, **kw)["sample"].transpose(1, 2)  # -> (N,T,C)
        return self.tensor2dtype(sample.contiguous(), return_dtype)  # a copy: the chain's buffers are reused

    @th.no_grad()
    def generate_sequence(self, wav_seqs, wav_sr, pose_dim, pose_fps, pose_window_len, pose_seed_len,
                          return_dtype="tensor", smooth_trans=True, trans_factor=None, init_poses=None,
                          sample_alg="ddim", batch_size=64, device="cpu", progress=True):
        """Long-form windowed generation (generator.py:80-195): windows are serial (each is in-painted from the
        previous window's tail), clips are the parallel axis.  Two upstream behaviours are kept on purpose:
        the audio window of division k>=1 is computed from `pose_start_frame` before it is advanced
        (generator.py:172-174), and `init_poses=None` is accepted here (upstream crashes on `.to`, :105)."""
        assert len(wav_seqs.shape) == 2, "Provide batch dimension"
        if init_poses is not None:
            assert len(init_poses.shape) == 3, "Provide batch dimension"
            assert len(init_poses) == len(wav_seqs), "Init pose batch size does not meet wav_seqs."
            init_poses = init_poses.to(device)
        wav_seqs = wav_seqs.to(device)
        num_seq = len(wav_seqs)
        num_batches = int(np.ceil(num_seq / batch_size))
        seq_len = wav_seqs.shape[1] // wav_sr * pose_fps
        wav_seq_len = wav_seqs.shape[1]
        stride = pose_window_len - pose_seed_len
        num_division = int(np.ceil(seq_len / stride))
        if (seq_len - pose_seed_len) % stride == 0:
            num_division -= 1
        wav_window_len = int(wav_sr * pose_window_len / pose_fps)
        outs = []
        for b in range(num_batches):
            wav_seq = wav_seqs[b * batch_size:(b + 1) * batch_size]
            init = None if init_poses is None else init_poses[b * batch_size:(b + 1) * batch_size]
            n = len(wav_seq)
            w0, w1, pose_start = 0, wav_window_len, 0
            samples, sample, inpaint_poses = [], None, None
            for k in range(num_division):
                wavs = wav_seq[:, w0:w1]
                masks = th.ones((n, pose_window_len, 1), device=device)
                masks[:, pose_seed_len:] = 0
                if k == 0:
                    if init is None:
                        inpaint_poses = masks = None
                    else:
                        inpaint_poses = th.zeros((n, pose_window_len, pose_dim), device=device)
                        inpaint_poses[:, :pose_seed_len] = init
                else:
                    if inpaint_poses is None:
                        inpaint_poses = th.zeros((n, pose_window_len, pose_dim), device=device)
                    inpaint_poses[:, :pose_seed_len] = sample[:, -pose_seed_len:]
                if w1 > wav_seq_len:
                    wavs = th.cat([wavs, th.zeros((n, w1 - wav_seq_len), device=device)], dim=1)
                sample = self.generate_sample((n, pose_dim, pose_window_len), wavs, inpaint_poses=inpaint_poses,
                                              inpaint_masks=masks, sample_alg=sample_alg, trans_factor=trans_factor,
                                              pose_seed_len=pose_seed_len, device=device, progress=progress)
                samples.append(sample)
                w0 = int(pose_start / pose_fps * wav_sr)
                w1 = w0 + wav_window_len
                pose_start += stride
            combined = []
            for i, x in enumerate(samples):
                if smooth_trans and i > 0:
                    ratio = th.arange(0, 1, 1 / pose_seed_len, device=device)[:pose_seed_len].view(1, -1, 1)
                    head = x[:, :pose_seed_len] * ratio + samples[i - 1][:, -pose_seed_len:] * (1 - ratio)
                    x = th.cat([head, x[:, pose_seed_len:]], dim=1)
                combined.append(x[:, :-pose_seed_len] if i < len(samples) - 1 else x)
            outs.append(th.cat(combined, dim=1)[:, :seq_len])
        return self.tensor2dtype(th.concat(outs, dim=0), return_dtype)

    @th.no_grad()
    def gpu_warm_up_ddim(self, shape, model_kwargs, device, num_iteration: int = 10):
        """generator.py:16-32: untimed DDIM chains (the first one also captures the chain graph and packs the weights)."""
        for _ in range(num_iteration):
            self.diffusion.ddim_sample_loop(self.model, shape, model_kwargs=model_kwargs, device=device, progress=False)

    @th.no_grad()
    def eval_infer_time_ddim(self, shape, model_kwargs, sample_alg="ddim", repetitions=10, device="cpu"):
        """generator.py:47-78: 10 warm-up chains, then `repetitions` timed with CUDA events -> (mean ms, std ms)."""
        sample_func = self._choose_sample_func(sample_alg)
        self.gpu_warm_up_ddim(shape, model_kwargs, device)
        if sample_alg != "ddim":  # the warm-up above is DDIM (as upstream); keep this algorithm's graph capture untimed too
            sample_func(self.model, shape, model_kwargs=model_kwargs, device=device, progress=False)
        timings = np.zeros((repetitions, 1))
        start, end = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
        for rep in range(repetitions):
            start.record()
            sample_func(self.model, shape, model_kwargs=model_kwargs, device=device, progress=False)
            end.record()
            th.cuda.synchronize()
            timings[rep] = start.elapsed_time(end)
        return np.sum(timings) / repetitions, np.std(timings)

    @th.no_grad()
    def eval_bpd(self, poses: th.Tensor, wavs: th.Tensor, pose_seed_len: int = None, noise_tape: th.Tensor = None):
        """generator.py:197-216: variational bound of `poses` (N,T,C) given `wavs`, on the device the model lives on."""
        device = next(self.model.parameters()).device
        poses = poses.to(device)
        model_kwargs = {"wav": wavs.to(device)}
        if getattr(self.model, "model_type", None) == "inpaint":
            assert pose_seed_len is not None, "Provide pose_seed_len for inpaint model."
            inpaint_masks = th.ones_like(poses)[:, :, :1]  # (N,T,1)
            inpaint_masks[:, pose_seed_len:] = 0
            model_kwargs["inpaint_pose"] = poses.clone().transpose(0, 1)  # (T,N,C)
            model_kwargs["inpaint_mask"] = inpaint_masks.transpose(0, 1)  # (T,N,1)
        return self.diffusion.calc_bpd_loop(self.model, x_start=poses.transpose(1, 2),  # -> (N,C,T)
                                            model_kwargs=model_kwargs, noise_tape=noise_tape)

    @staticmethod
    def tensor2dtype(x: th.Tensor, dtype: str):
        if dtype == "tensor":
            return x
        if dtype == "cpu_tensor":
            return x.cpu()
        if dtype == "array":
            return x.cpu().numpy()
        raise ValueError(f"Unsupported dtype: {dtype}")
