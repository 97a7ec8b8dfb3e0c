"""Diffusion process object: float64 schedule tables (bit-exact replay of the reference pipeline) and the
sampling loops, which hand the whole chain to the CUDA engine.

Mirrors `models/modules/gaussian_diffusion.py` (`GaussianDiffusion` :75-143, loops :331-412, :414-529),
`models/modules/respace.py` (`space_timesteps` :13-68, `GaussianSpacedDiffusion` :71-101) and
`models/model_creation.py:30-48` (`create_diffusion`).
"""
import math

import numpy as np
import torch as th


def get_named_beta_schedule(schedule_name, num_diffusion_timesteps):
    """gaussian_diffusion.py:20-61."""
    if schedule_name == "linear":
        scale = 1000 / num_diffusion_timesteps
        return np.linspace(scale * 0.0001, scale * 0.02, num_diffusion_timesteps, dtype=np.float64)
    if schedule_name == "squaredcos_cap_v2":
        bar = lambda t: math.cos(t * math.pi / 2) ** 2  # noqa: E731
        n = num_diffusion_timesteps
        return np.array([min(1 - bar((i + 1) / n) / bar(i / n), 0.999) for i in range(n)])
    raise NotImplementedError(f"unknown beta schedule: {schedule_name}")


def space_timesteps(num_timesteps, section_counts):
    """respace.py:13-68: which base timesteps a (possibly shortened) sampling process keeps."""
    if isinstance(section_counts, str):
        if section_counts.startswith("path:"):
            return set(np.load(section_counts[len("path:"):]))
        if section_counts.startswith("ddim"):
            want = int(section_counts[len("ddim"):])
            for stride in range(1, num_timesteps):
                if len(range(0, num_timesteps, stride)) == want:
                    return set(range(0, num_timesteps, stride))
            raise ValueError(f"cannot create exactly {num_timesteps} steps with an integer stride")
        if section_counts == "fast27":
            steps = space_timesteps(num_timesteps, "10,10,3,2,2")
            steps.remove(num_timesteps - 1)
            steps.add(num_timesteps - 3)
            return steps
        section_counts = [int(x) for x in section_counts.split(",")]
    size_per, extra = divmod(num_timesteps, len(section_counts))
    start, kept = 0, []
    for i, count in enumerate(section_counts):
        size = size_per + (1 if i < extra else 0)
        if size < count:
            raise ValueError(f"cannot divide section of {size} steps into {count}")
        stride = 1 if count <= 1 else (size - 1) / (count - 1)
        cur = 0.0
        for _ in range(count):
            kept.append(start + round(cur))
            cur += stride
        start += size
    return set(kept)


class GaussianDiffusion:
    """Schedule tables in float64 (gaussian_diffusion.py:87-143). Only `fixed_small` variance exists upstream."""

    def __init__(self, *, betas, model_var_type):
        if model_var_type != "fixed_small":
            raise ValueError(f"unsupported model_var_type {model_var_type}")
        self.model_var_type = model_var_type
        betas = np.array(betas, dtype=np.float64)
        self.betas = betas
        assert len(betas.shape) == 1, "betas must be 1-D"
        assert (betas > 0).all() and (betas <= 1).all()
        self.num_timesteps = int(betas.shape[0])
        self.alphas = 1.0 - betas
        self.alphas_cumprod = np.cumprod(self.alphas, axis=0)
        self.alphas_cumprod_prev = np.append(1.0, self.alphas_cumprod[:-1])
        self.alphas_cumprod_next = np.append(self.alphas_cumprod[1:], 0.0)
        self.sqrt_alphas_cumprod = np.sqrt(self.alphas_cumprod)
        self.sqrt_one_minus_alphas_cumprod = np.sqrt(1.0 - self.alphas_cumprod)
        self.log_one_minus_alphas_cumprod = np.log(1.0 - self.alphas_cumprod)
        self.sqrt_recip_alphas_cumprod = np.sqrt(1.0 / self.alphas_cumprod)
        self.sqrt_recipm1_alphas_cumprod = np.sqrt(1.0 / self.alphas_cumprod - 1.0)
        self.sqrt_recip_alphas = np.sqrt(1.0 / self.alphas)
        self.posterior_variance = betas * (1.0 - self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)
        self.posterior_log_variance_clipped = np.log(np.append(self.posterior_variance[1], self.posterior_variance[1:]))
        self.posterior_mean_coef1 = betas * np.sqrt(self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)
        self.posterior_mean_coef2 = (1.0 - self.alphas_cumprod_prev) * np.sqrt(self.alphas) / (1.0 - self.alphas_cumprod)


class InpaintBlend:
    """The reference's `denoise_fn` closure (generator.py:255-281) as data, so it can run inside the fused
    update epilogue:  x0 <- (1-f)*m*seed + f*m*x0 + (1-m)*x0  with f the `trans_factor` ramp over frames."""

    def __init__(self, seed_poses, masks, trans_factor, pose_seed_len, n_frames):
        # chain-owned copies: the engine refreshes these buffers in place when a cached plan is reused, which must never
        # write through to the caller's inpaint_poses / inpaint_masks tensors
        self.seed = seed_poses.float().contiguous().clone()            # (N, T, C)
        self.mask = masks.float().reshape(masks.shape[0], -1).contiguous().clone()  # (N, T)
        if trans_factor is not None:
            assert 0 <= trans_factor <= 1
            assert pose_seed_len is not None, "Provide pose_seed_len when using trans_factor."
            ramp = th.arange(trans_factor, 1, (1 - trans_factor) / pose_seed_len, device=self.seed.device)
            self.factor = th.cat([ramp, th.ones(n_frames - ramp.numel(), device=self.seed.device)]).float().contiguous()
        else:
            self.factor = th.zeros(n_frames, device=self.seed.device)

    def slice(self, lo, hi):
        """The blend of clips [lo, hi) (SplitChain: every sub-chain owns its slice)."""
        out = object.__new__(InpaintBlend)
        out.seed, out.mask, out.factor = self.seed[lo:hi].contiguous().clone(), self.mask[lo:hi].contiguous().clone(), self.factor.clone()
        return out

    def __call__(self, pred_x_start):  # (N, C, T) -> (N, C, T); torch form, used by the un-fused progressive API
        p = pred_x_start.transpose(1, 2)
        f, m = self.factor[None, :, None], self.mask[:, :, None]
        return ((1 - f) * m * self.seed + f * m * p + (1 - m) * p).transpose(1, 2)


class GaussianSpacedDiffusion(GaussianDiffusion):
    """respace.py:71-101: keeps `use_timesteps` of a base process; betas are re-derived from the base
    alphas_cumprod (this changes the last bit of most betas even when every step is kept)."""

    def __init__(self, use_timesteps, **kwargs):
        self.use_timesteps = set(use_timesteps)
        self.original_num_steps = len(kwargs["betas"])
        base = GaussianDiffusion(**kwargs)
        last, new_betas, self.timestep_map = 1.0, [], []
        for i, acp in enumerate(base.alphas_cumprod):
            if i in self.use_timesteps:
                new_betas.append(1 - acp / last)
                last = acp
                self.timestep_map.append(i)
        kwargs["betas"] = np.array(new_betas)
        super().__init__(**kwargs)

    # ------------------------------------------------------------------ coefficient tables for the kernels
    def step_tables(self, alg="ddpm", eta=0.0):
        """fp32 per-step coefficients [A, B, C1, C2, sigma] consumed by gd_ddpm_update / gd_linear_ddpm.
        float64 -> fp32 happens here exactly as in `_extract_into_tensor` (gaussian_diffusion.py:691).
        For DDIM (:443-484) the update x0*sqrt(abar_prev) + s*(A x - x0)/B + sigma*z with
        sigma = eta*sqrt((1-abar_prev)/(1-abar))*sqrt(1-abar/abar_prev), s = sqrt(1-abar_prev-sigma^2) is the same affine
        form with C1 = sqrt(abar_prev) - s/B, C2 = s*A/B (eta = 0, the reference's only call: sigma = 0)."""
        A = th.from_numpy(self.sqrt_recip_alphas_cumprod).float()
        B = th.from_numpy(self.sqrt_recipm1_alphas_cumprod).float()
        if alg == "ddpm":
            C1 = th.from_numpy(self.posterior_mean_coef1).float()
            C2 = th.from_numpy(self.posterior_mean_coef2).float()
            sigma = th.exp(0.5 * th.from_numpy(self.posterior_log_variance_clipped).float())
        elif alg == "ddim":
            a64, b64 = self.sqrt_recip_alphas_cumprod, self.sqrt_recipm1_alphas_cumprod
            ab, abp = self.alphas_cumprod, self.alphas_cumprod_prev
            sig64 = eta * np.sqrt((1.0 - abp) / (1.0 - ab)) * np.sqrt(1.0 - ab / abp)
            sp, sq = np.sqrt(abp), np.sqrt(1.0 - abp - sig64 ** 2)
            C1 = th.from_numpy(sp - sq / b64).float()
            C2 = th.from_numpy(sq * a64 / b64).float()
            sigma = th.from_numpy(sig64).float()
        else:
            raise ValueError(f"Unsupported sample algorithm: {alg}")
        return A, B, C1, C2, sigma

    # ------------------------------------------------------------------ sampling loops (gaussian_diffusion.py:331-529)
    def _run(self, alg, model, shape, noise, denoise_fn, model_kwargs, device, progress, noise_tape, progressive, eta=0.0):
        from .engine import chain_for  # local import: the engine needs the CUDA library
        if device is None:
            device = next(model.parameters()).device
        assert isinstance(shape, (tuple, list))
        chain = chain_for(model, self, tuple(shape), alg, device, allow_split=not progressive, eta=float(eta))
        wav = (model_kwargs or {}).get("wav")
        if wav is None:
            raise ValueError("model_kwargs['wav'] is required")
        if noise is None:
            noise = th.randn(*shape, device=device)
        offset = model.input_offset(model_kwargs) if hasattr(model, "input_offset") else None
        chain.begin(noise.to(device), wav.to(device), denoise_fn=denoise_fn, noise_tape=noise_tape,
                    need_tape=(alg == "ddpm" or eta != 0.0), input_offset=offset)
        if progressive:
            return chain.iterate(progress)
        return chain.run(progress)

    def p_sample_loop(self, model, shape, model_kwargs, noise=None, denoise_fn=None, device=None, progress=False,
                      noise_tape=None):
        """Full ancestral chain as CUDA-graph replays; returns the last step's dict (keys as the reference's
        p_sample: sample, eps, pred_x_start).  `noise_tape` (n_steps, N, C, T), indexed by loop position
        k = 0..n-1 (first draw first), replaces the per-step `randn_like` draws."""
        return self._run("ddpm", model, shape, noise, denoise_fn, model_kwargs, device, progress, noise_tape, False)

    def p_sample_loop_progressive(self, model, shape, noise=None, model_kwargs=None, denoise_fn=None, device=None,
                                  progress=False, noise_tape=None):
        return self._run("ddpm", model, shape, noise, denoise_fn, model_kwargs, device, progress, noise_tape, True)

    def ddim_sample_loop(self, model, shape, noise=None, denoise_fn=None, model_kwargs=None, device=None,
                         progress=False, eta=0.0, noise_tape=None):
        """gaussian_diffusion.py:486-529.  eta != 0 adds sigma_t * noise per step (`noise_tape` replaces the randn_like draws)."""
        return self._run("ddim", model, shape, noise, denoise_fn, model_kwargs, device, progress, noise_tape, False, eta=eta)

    def ddim_sample_loop_progressive(self, model, shape, noise=None, denoise_fn=None, model_kwargs=None, device=None,
                                     progress=False, eta=0.0, noise_tape=None):
        return self._run("ddim", model, shape, noise, denoise_fn, model_kwargs, device, progress, noise_tape, True, eta=eta)


    # ------------------------------------------------------------------ variational bound (gaussian_diffusion.py:571-678)
    @th.no_grad()
    def calc_bpd_loop(self, model, x_start, model_kwargs, noise_tape=None):
        """Variational lower bound in bits/dim, term by term (`calc_bpd_loop` :635-678, `_vb_terms_bpd` :571-610,
        `_prior_bpd` :612-633).  Every timestep is an independent denoiser call on x_t = q_sample(x_start, t, noise)
        (:182-205): the engine's step is teacher-forced with x_t and its fused final-projection epilogue hands back
        pred_x_start; the bound terms are elementwise maps and one mean per clip, done with the reference's own op
        order in fp32.  `noise_tape` (n_steps, N, C, T) in loop order replaces the per-step randn_like draws."""
        from .engine import chain_for  # local import: the engine needs the CUDA library
        device = x_start.device
        x_start = x_start.float().contiguous()
        shape = tuple(x_start.shape)
        wav = (model_kwargs or {}).get("wav")
        if wav is None:
            raise ValueError("model_kwargs['wav'] is required")
        chain = chain_for(model, self, shape, "ddpm", device)
        offset = model.input_offset(model_kwargs) if hasattr(model, "input_offset") else None
        chain.begin(th.zeros(shape, device=device), wav.to(device), need_tape=False, input_offset=offset)
        n = self.num_timesteps
        tab = lambda a: th.from_numpy(a).to(device).float()  # noqa: E731  (float64 table -> fp32 at gather, :691)
        sa, s1a = tab(self.sqrt_alphas_cumprod), tab(self.sqrt_one_minus_alphas_cumprod)
        c1, c2 = tab(self.posterior_mean_coef1), tab(self.posterior_mean_coef2)
        lv = tab(self.posterior_log_variance_clipped)
        ra, rm1 = tab(self.sqrt_recip_alphas_cumprod), tab(self.sqrt_recipm1_alphas_cumprod)
        flat = lambda z: z.mean(dim=list(range(1, z.dim())))  # noqa: E731
        ln2 = float(np.log(2.0))
        vb, x_start_mse, mse = [], [], []
        for k, t in enumerate(range(n - 1, -1, -1)):
            noise = noise_tape[k].to(device).float() if noise_tape is not None else th.randn_like(x_start)
            x_t = sa[t] * x_start + s1a[t] * noise
            chain.set_state(x_t, t)
            chain.step_eager()
            pred = chain.x0.clone()  # pred_x_start of p_mean_variance
            true_mean = c1[t] * x_start + c2[t] * x_t
            mean = c1[t] * pred + c2[t] * x_t
            kl = 0.5 * (-1.0 + lv[t] - lv[t] + th.exp(lv[t] - lv[t]) + ((true_mean - mean) ** 2) * th.exp(-lv[t]))
            kl = flat(kl) / ln2
            centered = (x_start - mean) * th.exp(-(0.5 * lv[t]))
            log_probs = (-centered ** 2 / 2) - th.log(th.sqrt(2 * th.tensor(math.pi, device=device)))
            decoder_nll = flat(-log_probs) / ln2
            vb.append(decoder_nll if t == 0 else kl)
            x_start_mse.append(flat((pred - x_start) ** 2))
            eps = (ra[t] * x_t - pred) / rm1[t]
            mse.append(flat((eps - noise) ** 2))
        vb, x_start_mse, mse = th.stack(vb, dim=1), th.stack(x_start_mse, dim=1), th.stack(mse, dim=1)
        qt_mean = sa[n - 1] * x_start
        qt_logvar = tab(self.log_one_minus_alphas_cumprod)[n - 1]
        kl_prior = 0.5 * (-1.0 + 0.0 - qt_logvar + th.exp(qt_logvar - 0.0) + ((qt_mean - 0.0) ** 2) * math.exp(-0.0))
        prior_bpd = flat(kl_prior) / ln2
        return {"total_bpd": vb.sum(dim=1) + prior_bpd, "prior_bpd": prior_bpd, "x_start_mse": x_start_mse, "vb": vb,
                "mse": mse}


def create_diffusion(diffusion_params, is_training):
    """model_creation.py:30-48."""
    if diffusion_params.type != "gaussian":
        raise ValueError
    betas = get_named_beta_schedule(diffusion_params.noise_schedule, diffusion_params.diffusion_steps)
    if not diffusion_params.timestep_respacing or is_training:
        respacing = [diffusion_params.diffusion_steps]
    else:
        respacing = diffusion_params.timestep_respacing
    return GaussianSpacedDiffusion(use_timesteps=space_timesteps(diffusion_params.diffusion_steps, respacing),
                                   betas=betas, model_var_type=diffusion_params.model_var_type)
