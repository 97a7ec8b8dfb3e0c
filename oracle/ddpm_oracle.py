"""CPU oracle for the DDPM reverse-sampling hot path — TEST INFRASTRUCTURE ONLY.

A functional fp32 (tables: float64 numpy) restatement of the reference's algorithm, written from
the reference's behaviour (file:line citations are to the reference repository).  It works
directly on a reference-format ``state_dict`` (dict name -> tensor) and never touches the
package's CUDA path.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU
baseline legs may import it; the product must not.

Parity status: PINNED.  ``tests/golden/make_golden.py`` runs the real reference (imported from
/root/reference in the build container) on deterministic synthetic weights / speech / noise and
commits its schedule tables, conditioning features, per-step eps and pose checkpoints;
``tests/test_oracle_golden.py`` checks this file against those vectors.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------- schedule tables
def linear_betas(num_steps):
    """get_named_beta_schedule('linear') — gaussian_diffusion.py:20-33."""
    scale = 1000 / num_steps
    return np.linspace(scale * 0.0001, scale * 0.02, num_steps, dtype=np.float64)


def cosine_betas(num_steps, max_beta=0.999):
    """'squaredcos_cap_v2' — gaussian_diffusion.py:34-61."""
    ab = lambda t: math.cos(t * math.pi / 2) ** 2  # noqa: E731
    return np.array([min(1 - ab((i + 1) / num_steps) / ab(i / num_steps), max_beta) for i in range(num_steps)])


def space_timesteps(num_timesteps, section_counts):
    """respace.py:13-68 (the 'path:' form is not restated)."""
    if isinstance(section_counts, str):
        if section_counts.startswith("ddim"):
            want = int(section_counts[4:])
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
    start, out = 0, []
    for i, count in enumerate(section_counts):
        size = size_per + (1 if i < extra else 0)
        if size < count:
            raise ValueError(f"cannot divide section of {size} steps into {count}")
        stride = 1 if count <= 1 else (size - 1) / (count - 1)
        cur = 0.0
        for _ in range(count):
            out.append(start + round(cur))
            cur += stride
        start += size
    return set(out)


def diffusion_tables(betas):
    """GaussianDiffusion.__init__ — gaussian_diffusion.py:95-143, float64 throughout."""
    betas = np.array(betas, dtype=np.float64)
    alphas = 1.0 - betas
    ac = np.cumprod(alphas, axis=0)
    ac_prev = np.append(1.0, ac[:-1])
    post_var = betas * (1.0 - ac_prev) / (1.0 - ac)
    return {
        "betas": betas,
        "alphas_cumprod": ac,
        "alphas_cumprod_prev": ac_prev,
        "sqrt_alphas_cumprod": np.sqrt(ac),
        "sqrt_one_minus_alphas_cumprod": np.sqrt(1.0 - ac),
        "log_one_minus_alphas_cumprod": np.log(1.0 - ac),
        "sqrt_recip_alphas_cumprod": np.sqrt(1.0 / ac),
        "sqrt_recipm1_alphas_cumprod": np.sqrt(1.0 / ac - 1.0),
        "posterior_variance": post_var,
        "posterior_log_variance_clipped": np.log(np.append(post_var[1], post_var[1:])),
        "posterior_mean_coef1": betas * np.sqrt(ac_prev) / (1.0 - ac),
        "posterior_mean_coef2": (1.0 - ac_prev) * np.sqrt(alphas) / (1.0 - ac),
    }


def spaced_diffusion_tables(noise_schedule="linear", diffusion_steps=1000, timestep_respacing=""):
    """create_diffusion (model_creation.py:30-48) + GaussianSpacedDiffusion.__init__ (respace.py:80-93):
    even with no respacing the betas are RE-DERIVED as 1 - acp_i/acp_{i-1} from the base process."""
    base = linear_betas(diffusion_steps) if noise_schedule == "linear" else cosine_betas(diffusion_steps)
    use = space_timesteps(diffusion_steps, timestep_respacing if timestep_respacing else [diffusion_steps])
    base_ac = np.cumprod(1.0 - np.array(base, dtype=np.float64), axis=0)
    last, new_betas, tmap = 1.0, [], []
    for i, a in enumerate(base_ac):
        if i in use:
            new_betas.append(1 - a / last)
            last = a
            tmap.append(i)
    tabs = diffusion_tables(np.array(new_betas))
    tabs["timestep_map"] = np.array(tmap, dtype=np.int64)
    return tabs


# ----------------------------------------------------------------------------- small pieces
def timestep_embedding(t, dim, max_period=10000):
    """diffusion_step_embedding — nn.py:17-35 (cos first, then sin)."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32) / half)
    args = t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def step_token(sd, t, d_model):
    """DiffusionStepEncoder — nn.py:38-52."""
    e = timestep_embedding(t, d_model)
    h = F.silu(F.linear(e, sd["diffusion_step_encoder.proj.0.weight"], sd["diffusion_step_encoder.proj.0.bias"]))
    return F.linear(h, sd["diffusion_step_encoder.proj.2.weight"], sd["diffusion_step_encoder.proj.2.bias"])


def positional_encoding(d_model, length):
    """get_positional_encoding — transformer.py:157-166 (sin at even, cos at odd features)."""
    pe = torch.zeros(length, d_model)
    pos = torch.arange(0, length, dtype=torch.float32).unsqueeze(1)
    div = torch.exp(torch.arange(0, d_model, 2, dtype=torch.float32) * -(math.log(10000.0) / d_model))
    pe[:, 0::2] = torch.sin(pos * div)
    pe[:, 1::2] = torch.cos(pos * div)
    return pe


def _ln(sd, key, x):
    return F.layer_norm(x, (x.shape[-1],), sd[key + ".weight"], sd[key + ".bias"], 1e-5)


def _dconv(sd, key, u, heads):
    """SpatialDepthWiseConv — transformer.py:19-44: k=3 depth-wise FIR over tokens, zero 'same' padding,
    taps shared by all heads.  u: (N, L, d)."""
    n, L, d = u.shape
    dk = d // heads
    z = u.view(n, L, heads, dk).permute(0, 2, 3, 1).reshape(n * heads, dk, L)
    z = F.conv1d(z, sd[key + ".1.conv.weight"], sd[key + ".1.conv.bias"], padding=1, groups=dk)
    return z.view(n, heads, dk, L).permute(0, 3, 1, 2)  # (N, L, H, dk)


def mdha(sd, key, q_in, kv_in, heads):
    """MultiDConvHeadAttention — transformer.py:88-126. q_in (N,Lq,d), kv_in (N,Lk,d)."""
    d = q_in.shape[-1]
    q = _dconv(sd, key + ".query", F.linear(q_in, sd[key + ".query.0.linear.weight"], sd[key + ".query.0.linear.bias"]), heads)
    k = _dconv(sd, key + ".key", F.linear(kv_in, sd[key + ".key.0.linear.weight"], sd[key + ".key.0.linear.bias"]), heads)
    v = _dconv(sd, key + ".value", F.linear(kv_in, sd[key + ".value.0.linear.weight"], sd[key + ".value.0.linear.bias"]), heads)
    s = torch.einsum("nihd,njhd->nhij", q, k) * (1.0 / math.sqrt(d // heads))
    p = torch.softmax(s, dim=-1)  # softmax over keys (transformer.py:72,113)
    o = torch.einsum("nhij,njhd->nihd", p, v).reshape(q_in.shape[0], q_in.shape[1], d)
    return F.linear(o, sd[key + ".output.weight"], sd[key + ".output.bias"])


def ffn(sd, key, x):
    """FeedForward with SquaredReLU — transformer.py:8-16,129-154."""
    h = torch.relu(F.linear(x, sd[key + ".layer1.weight"], sd[key + ".layer1.bias"]))
    return F.linear(h * h, sd[key + ".layer2.weight"], sd[key + ".layer2.bias"])


def _n_layers(sd):
    return 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("pose_decoder.layers."))


# ----------------------------------------------------------------------------- speech encoder
def _hz_to_mel(f):
    return 2595.0 * math.log10(1.0 + f / 700.0)


def mel_filterbank(n_freqs=513, f_min=0.0, f_max=8000.0, n_mels=128, sample_rate=16000):
    """torchaudio.functional.melscale_fbanks(norm=None, mel_scale='htk') restated."""
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_pts = torch.linspace(_hz_to_mel(f_min), _hz_to_mel(f_max), n_mels + 2)
    f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return torch.max(torch.zeros(1), torch.min(down, up))


def mel_frontend(wav, fb=None, window=None):
    """PreEmphasis (ha2g/model/utils.py:22-37) -> MelSpectrogram(16 kHz, n_fft 1024, hop 512, 128 mels,
    power 2, centre/reflect) -> +1e-6 -> InstanceNorm1d(128)  (speech_encoder.py:18-26,50-51)."""
    x = F.pad(wav.unsqueeze(1), (1, 0), "reflect")
    x = F.conv1d(x, torch.tensor([[[-0.97, 1.0]]], dtype=wav.dtype)).squeeze(1)
    window = torch.hann_window(1024) if window is None else window
    spec = torch.stft(x, n_fft=1024, hop_length=512, win_length=1024, window=window, center=True, pad_mode="reflect",
                      normalized=False, onesided=True, return_complex=True).abs().pow(2.0)
    fb = mel_filterbank() if fb is None else fb
    mel = torch.matmul(spec.transpose(-1, -2), fb).transpose(-1, -2) + 1e-6
    return F.instance_norm(mel, eps=1e-5)


def _bn(sd, key, x):
    return F.batch_norm(x, sd[key + ".running_mean"], sd[key + ".running_var"], sd[key + ".weight"], sd[key + ".bias"],
                        False, 0.1, 1e-5)


def _se_block(sd, key, x, stride):
    """SEBasicBlock.forward — ResNetBlocks.py:22-37 (note relu BEFORE bn1), SELayer :81-96."""
    out = _bn(sd, key + ".bn1", torch.relu(F.conv2d(x, sd[key + ".conv1.weight"], None, stride, 1)))
    out = _bn(sd, key + ".bn2", F.conv2d(out, sd[key + ".conv2.weight"], None, 1, 1))
    y = out.mean(dim=(2, 3))
    y = torch.relu(F.linear(y, sd[key + ".se.fc.0.weight"], sd[key + ".se.fc.0.bias"]))
    y = torch.sigmoid(F.linear(y, sd[key + ".se.fc.2.weight"], sd[key + ".se.fc.2.bias"]))
    out = out * y[:, :, None, None]
    res = x
    if key + ".downsample.0.weight" in sd:
        res = _bn(sd, key + ".downsample.1", F.conv2d(x, sd[key + ".downsample.0.weight"], None, stride, 0))
    return torch.relu(out + res)


def _pyramid_head(sd, name, feat, shuffle):
    """conv_{low,mid,high} -> relu -> bn -> flatten (C*H) per time frame -> fc  (ResNetSE34V2.py:156-189)."""
    p = "speech_encoder.wav_encoder.feat_extractor."
    if shuffle > 1:
        feat = F.pixel_shuffle(feat, shuffle)
    f = _bn(sd, p + "bn_" + name, torch.relu(F.conv2d(feat, sd[p + f"conv_{name}.weight"], sd[p + f"conv_{name}.bias"])))
    n = f.shape[0]
    f = f.reshape(n, -1, f.shape[-1]).transpose(1, 2)
    return F.linear(f, sd[p + f"fc_{name}.weight"], sd[p + f"fc_{name}.bias"])  # (N, T_k, 32)


def speech_features(sd, wav):
    """HA2GSpeechEncoder.forward — speech_encoder.py:37-61: three (N, T_k, d_model) feature pyramids."""
    p = "speech_encoder.wav_encoder.feat_extractor."
    mel = mel_frontend(wav, sd.get("speech_encoder.wav2spec.1.mel_scale.fb"),
                       sd.get("speech_encoder.wav2spec.1.spectrogram.window"))
    x = mel.unsqueeze(1)
    x = _bn(sd, p + "bn1", torch.relu(F.conv2d(x, sd[p + "conv1.weight"], sd[p + "conv1.bias"], 1, 1)))
    feats = []
    for li, (nblk, stride) in enumerate(((3, 1), (4, 2), (6, 2), (3, 2)), start=1):
        for b in range(nblk):
            x = _se_block(sd, f"{p}layer{li}.{b}", x, stride if b == 0 else 1)
        feats.append(x)
    proj = lambda f: F.linear(f, sd["speech_encoder.wav_proj_layer.weight"], sd["speech_encoder.wav_proj_layer.bias"])  # noqa: E731
    return (proj(_pyramid_head(sd, "low", feats[1], 1)), proj(_pyramid_head(sd, "mid", feats[2], 2)),
            proj(_pyramid_head(sd, "high", feats[3], 4)))


# ----------------------------------------------------------------------------- denoisers
def inpaint_offset(sd, inpaint_pose, inpaint_mask):
    """Speech2GestureModelInpaint.myforward model.py:155-165: proj([pose*mask | mask]) with proj = Linear, SiLU, Linear,
    SiLU, Linear.  inpaint_pose (N,T,C), inpaint_mask (N,T,1) -> (N,C,T) offset of the denoiser input."""
    h = torch.cat([inpaint_pose * inpaint_mask, inpaint_mask], dim=-1)
    h = F.silu(F.linear(h, sd["proj.0.weight"], sd["proj.0.bias"]))
    h = F.silu(F.linear(h, sd["proj.2.weight"], sd["proj.2.bias"]))
    return F.linear(h, sd["proj.4.weight"], sd["proj.4.bias"]).permute(0, 2, 1)


def denoiser(sd, model_type, heads, x_t, t, feats, offset=None):
    """model(x_t (N,C,T), t (N,)) -> eps (N,C,T) given the (loop-invariant) speech features.
    tedexp: Speech2GestureModel.myforward model.py:41-73 + CrossAttention nn.py:428-447,90-125
    beat:   Speech2GestureModelV2.myforward model.py:81-117 + OnewayCrossAttention nn.py:216-228,154-174
    inpaint: the tedexp wrapper applied to x_t + offset (inpaint_offset above)"""
    if model_type == "inpaint":
        x_t, model_type = x_t + offset, "default"
    d = sd["pose_decoder.emb_x.weight"].shape[0]
    z_low, z_mid, z_high = feats
    zt = step_token(sd, t, d).unsqueeze(1)  # (N,1,d)
    x = x_t.permute(0, 2, 1)  # (N,T,C)
    L = _n_layers(sd)
    lin = lambda key, v: F.linear(v, sd[key + ".weight"], sd[key + ".bias"])  # noqa: E731
    pd = "pose_decoder."
    if model_type == "default":
        mem = torch.cat([zt, z_low, z_mid, z_high], dim=1)
        X, M = lin(pd + "emb_x", x), lin(pd + "emb_mem", mem)
        Tx = X.shape[1]
        H = torch.cat([X, M], dim=1)
        H = H + positional_encoding(d, H.shape[1])
        X, M = H[:, :Tx], H[:, Tx:]
        for l in range(L):
            k = f"{pd}layers.{l}."
            z = _ln(sd, k + "norm_self_attn", X)
            X = X + mdha(sd, k + "self_attn", z, z, heads)
            z = _ln(sd, k + "norm_self_attn_mem", M)
            M = M + mdha(sd, k + "self_attn_mem", z, z, heads)
            H = torch.cat([X, M], dim=1)
            z = _ln(sd, k + "norm_cross_attn", H)
            H = H + mdha(sd, k + "cross_attn", z, z, heads)
            X, M = H[:, :Tx], H[:, Tx:]
            X = X + ffn(sd, k + "feed_forward", _ln(sd, k + "norm_ff", X))
            if k + "feed_forward_mem.layer1.weight" in sd:
                M = M + ffn(sd, k + "feed_forward_mem", _ln(sd, k + "norm_ff_mem", M))
    elif model_type == "s2g_v2":
        longest = max(z_low.shape[1], z_mid.shape[1], z_high.shape[1])
        padf = lambda z: F.pad(z, (0, 0, longest - z.shape[1], 0))  # zero rows PREPENDED (model.py:97-103)  # noqa: E731
        z = lin("blend_layer", torch.cat([padf(z_low), padf(z_mid), padf(z_high)], dim=-1))
        mem = torch.cat([zt, z], dim=1)
        X = lin(pd + "emb_x", x)
        X = X + positional_encoding(d, X.shape[1])
        M = lin(pd + "emb_mem", mem)
        M = M + positional_encoding(d, M.shape[1])
        for l in range(L):
            k = f"{pd}layers.{l}."
            zq = _ln(sd, k + "norm_self_attn", X)
            X = X + mdha(sd, k + "self_attn", zq, zq, heads)
            X = X + mdha(sd, k + "cross_attn", _ln(sd, k + "norm_cross_attn", X), M, heads)
            X = X + ffn(sd, k + "feed_forward", _ln(sd, k + "norm_ff", X))
    else:
        raise ValueError(f"oracle: unsupported model type {model_type}")
    out = lin(pd + "out_layers.1", _ln(sd, pd + "out_layers.0", X))
    return out.permute(0, 2, 1)


# ----------------------------------------------------------------------------- sampler
def _f32(tab, i):
    return torch.tensor(tab[i]).float()  # float64 -> fp32 at gather time (gaussian_diffusion.py:691)


def inpaint_blend(x0, seed, mask, factor):
    """denoise_fn — generator.py:265-280. x0 (N,C,T); seed (N,T,C); mask (N,T,1); factor (1,T,1) or 0."""
    p = x0.transpose(1, 2)
    p = (1 - factor) * mask * seed + factor * mask * p + (1 - mask) * p
    return p.transpose(1, 2)


def transition_factor(trans_factor, pose_seed_len, T):
    """generator.py:258-268."""
    if trans_factor is None:
        return 0
    f = torch.arange(trans_factor, 1, (1 - trans_factor) / pose_seed_len)[None, :, None]
    return torch.cat([f, torch.ones((1, T - f.size(1), 1))], dim=1)


def ddpm_step(tabs, i, x, eps, z, blend=None):
    """p_mean_variance + p_sample for loop index i — gaussian_diffusion.py:287-292,215-220,326-328."""
    x0 = _f32(tabs["sqrt_recip_alphas_cumprod"], i) * x - _f32(tabs["sqrt_recipm1_alphas_cumprod"], i) * eps
    if blend is not None:
        x0 = blend(x0)
    mean = _f32(tabs["posterior_mean_coef1"], i) * x0 + _f32(tabs["posterior_mean_coef2"], i) * x
    nz = 0.0 if i == 0 else 1.0
    return mean + nz * torch.exp(0.5 * _f32(tabs["posterior_log_variance_clipped"], i)) * z, x0


def ddim_step(tabs, i, x, eps, blend=None, eta=0.0, z=None):
    """ddim_sample — gaussian_diffusion.py:443-484 (the reference only ever runs eta = 0; `z` is the step's randn_like draw)."""
    a, b = _f32(tabs["sqrt_recip_alphas_cumprod"], i), _f32(tabs["sqrt_recipm1_alphas_cumprod"], i)
    x0 = a * x - b * eps
    if blend is not None:
        x0 = blend(x0)
    eps2 = (a * x - x0) / b
    ab, ab_prev = _f32(tabs["alphas_cumprod"], i), _f32(tabs["alphas_cumprod_prev"], i)
    sigma = eta * torch.sqrt((1 - ab_prev) / (1 - ab)) * torch.sqrt(1 - ab / ab_prev)
    mean_pred = x0 * torch.sqrt(ab_prev) + torch.sqrt(1 - ab_prev - sigma ** 2) * eps2
    if eta != 0.0 and i != 0 and z is not None:
        mean_pred = mean_pred + sigma * z
    return mean_pred, x0


@torch.no_grad()
def sample_chain(sd, model_type, heads, tabs, x_T, wav, tape, alg="ddpm", steps=None, blend=None,
                 reencode_every_step=False, record=None, offset=None, eta=0.0):
    """p_sample_loop / ddim_sample_loop — gaussian_diffusion.py:368-412,486-529.
    tape[k] is the k-th randn_like draw (loop order i = n-1 .. 0); `steps` restricts to the first
    `steps` iterations (bounded CPU baseline).  reencode_every_step=True reproduces the reference as
    shipped, which reruns the speech encoder inside every denoiser call (model.py:54-56)."""
    n = len(tabs["betas"])
    x = x_T
    feats = None if reencode_every_step else speech_features(sd, wav)
    for k, i in enumerate(range(n - 1, -1, -1)):
        if steps is not None and k >= steps:
            break
        t = torch.full((x.shape[0],), int(tabs["timestep_map"][i]), dtype=torch.long)
        f = speech_features(sd, wav) if reencode_every_step else feats
        eps = denoiser(sd, model_type, heads, x, t, f, offset=offset)
        if alg == "ddpm":
            x_next, x0 = ddpm_step(tabs, i, x, eps, tape[k] if tape is not None else torch.zeros_like(x), blend)
        else:
            x_next, x0 = ddim_step(tabs, i, x, eps, blend, eta=eta, z=tape[k] if tape is not None else None)
        if record is not None:
            record(i, x, eps, x_next)
        x = x_next
    return x


@torch.no_grad()
def bpd_loop(sd, model_type, heads, tabs, x_start, wav, tape):
    """calc_bpd_loop + _vb_terms_bpd + _prior_bpd - gaussian_diffusion.py:571-678 with losses.py:6-56.
    x_start (N,C,T); tape[k] is the k-th randn_like draw (loop order t = n-1 .. 0).  fixed_small variance:
    the model log-variance is posterior_log_variance_clipped, as is the true posterior's."""
    n = len(tabs["betas"])
    feats = speech_features(sd, wav)
    flat = lambda z: z.mean(dim=list(range(1, z.dim())))  # noqa: E731
    ln2 = float(np.log(2.0))
    vb, x0_mse, mse = [], [], []
    for k, t in enumerate(range(n - 1, -1, -1)):
        noise = tape[k]
        x_t = _f32(tabs["sqrt_alphas_cumprod"], t) * x_start + _f32(tabs["sqrt_one_minus_alphas_cumprod"], t) * noise
        tt = torch.full((x_start.shape[0],), int(tabs["timestep_map"][t]), dtype=torch.long)
        eps_model = denoiser(sd, model_type, heads, x_t, tt, feats)
        a, b = _f32(tabs["sqrt_recip_alphas_cumprod"], t), _f32(tabs["sqrt_recipm1_alphas_cumprod"], t)
        pred = a * x_t - b * eps_model
        c1, c2 = _f32(tabs["posterior_mean_coef1"], t), _f32(tabs["posterior_mean_coef2"], t)
        lv = _f32(tabs["posterior_log_variance_clipped"], t)
        true_mean, mean = c1 * x_start + c2 * x_t, c1 * pred + c2 * x_t
        kl = 0.5 * (-1.0 + lv - lv + torch.exp(lv - lv) + ((true_mean - mean) ** 2) * torch.exp(-lv))
        centered = (x_start - mean) * torch.exp(-(0.5 * lv))
        nll = -((-centered ** 2 / 2) - torch.log(torch.sqrt(2 * torch.tensor(math.pi))))
        vb.append(flat(nll) / ln2 if t == 0 else flat(kl) / ln2)
        x0_mse.append(flat((pred - x_start) ** 2))
        mse.append(flat(((a * x_t - pred) / b - noise) ** 2))
    vb, x0_mse, mse = torch.stack(vb, 1), torch.stack(x0_mse, 1), torch.stack(mse, 1)
    qt_mean = _f32(tabs["sqrt_alphas_cumprod"], n - 1) * x_start
    qt_lv = _f32(tabs["log_one_minus_alphas_cumprod"], n - 1)
    prior = flat(0.5 * (-1.0 + 0.0 - qt_lv + torch.exp(qt_lv - 0.0) + (qt_mean ** 2) * 1.0)) / ln2
    return {"total_bpd": vb.sum(1) + prior, "prior_bpd": prior, "x_start_mse": x0_mse, "vb": vb, "mse": mse}
