"""Host-side driver of the sampling chain: packs weights once, hoists the loop-invariant conditioning, builds the
per-step launch plan over the C-ABI kernels (include/gd_b200.h) and replays it as a CUDA graph.

Data layout in HBM (per engine, N clips):
  x        (N, C, T) fp32      the sample, boundary layout of the reference (models/model.py:12-15)
  xa       [N*T, 128] bf16     x transposed to token rows, d_pose zero-padded to 128 (emb_x GEMM operand)
  H        [rows, d] fp32      residual stream(s); tedexp: pose rows [0, N*Tx) then memory rows [N*Tx, N*(Tx+Tm))
  xn/ao    [rows, d] bf16      LayerNorm output / attention output (GEMM operands)
  qkv      [rows, 3d] bf16     fused Q|K|V projection (fp32 in the `fp32act` precision mode)
  hid      [rows, 4d] bf16     FFN hidden after squared ReLU
  tape     (n_steps, N, C, T)  pre-generated Gaussian noise indexed by loop index i
The step index lives in a device int; every kernel that depends on it reads it there, so ONE captured
step graph serves all steps; by default the whole T-step chain is captured as one CUDA graph (`graph_steps` = 0), a
positive `graph_steps` captures that many consecutive steps per graph instead.
"""
import ctypes as C
import math
import os
import weakref

import torch as th

from . import _lib as gd
from .modules import positional_table, step_embedding_table

_POSE_PAD = 128


def _p(t):
    return None if t is None else t.data_ptr()


class PackedWeights:
    """bf16/fp32 repack of a denoiser's parameters for the kernels (done once per weight version)."""

    def __init__(self, model, diffusion, device):
        self.device = device
        # repacked on the HOST and uploaded once (a few hundred memcpys instead of ~1000 ATen cast / cat / fill launches)
        sd = {k: v.detach().to(device="cpu", dtype=th.float32) for k, v in model.state_dict().items()
              if v.dtype.is_floating_point}
        self.kind = model.pose_decoder.kind
        self.d = d = model.d_model
        self.C = model.d_pose
        self.heads = model.pose_decoder.heads
        self.n_layers = model.pose_decoder.n_layers
        bf = lambda w: w.to(th.bfloat16).contiguous().to(device)  # noqa: E731
        f32 = lambda w: w.float().contiguous().to(device)  # noqa: E731
        pd = "pose_decoder."
        w = sd[pd + "emb_x.weight"]
        wx = th.zeros(d, _POSE_PAD)
        wx[:, :self.C] = w
        self.embx_w, self.embx_b = bf(wx), f32(sd[pd + "emb_x.bias"])
        self.embm_w, self.embm_b = bf(sd[pd + "emb_mem.weight"]), f32(sd[pd + "emb_mem.bias"])
        wo = th.zeros(_POSE_PAD, d)
        wo[:self.C] = sd[pd + "out_layers.1.weight"]
        bo = th.zeros(_POSE_PAD)
        bo[:self.C] = sd[pd + "out_layers.1.bias"]
        self.out_w, self.out_b = bf(wo), f32(bo)
        self.out_ln = (f32(sd[pd + "out_layers.0.weight"]), f32(sd[pd + "out_layers.0.bias"]))

        def attn(prefix):
            a = {}
            names = ("query", "key", "value")
            a["wqkv"] = bf(th.cat([sd[f"{prefix}.{n}.0.linear.weight"] for n in names], 0))
            a["bqkv"] = f32(th.cat([sd[f"{prefix}.{n}.0.linear.bias"] for n in names], 0))
            a["taps"] = [f32(sd[f"{prefix}.{n}.1.conv.{p}"].reshape(-1, 3) if p == "weight" else sd[f"{prefix}.{n}.1.conv.{p}"])
                         for n in names for p in ("weight", "bias")]  # wq,bq,wk,bk,wv,bv
            a["wo"], a["bo"] = bf(sd[prefix + ".output.weight"]), f32(sd[prefix + ".output.bias"])
            return a

        def ffn(prefix):
            return {"w1": bf(sd[prefix + ".layer1.weight"]), "b1": f32(sd[prefix + ".layer1.bias"]),
                    "w2": bf(sd[prefix + ".layer2.weight"]), "b2": f32(sd[prefix + ".layer2.bias"])}

        def ln(prefix):
            return (f32(sd[prefix + ".weight"]), f32(sd[prefix + ".bias"]))

        self.layers = []
        for l in range(self.n_layers):
            k = f"{pd}layers.{l}."
            L = {"ln_sa": ln(k + "norm_self_attn"), "sa": attn(k + "self_attn"), "ln_ca": ln(k + "norm_cross_attn"),
                 "ca": attn(k + "cross_attn"), "ln_ff": ln(k + "norm_ff"), "ff": ffn(k + "feed_forward")}
            if self.kind == "cross_attention":
                L["ln_sam"], L["sam"] = ln(k + "norm_self_attn_mem"), attn(k + "self_attn_mem")
                if k + "feed_forward_mem.layer1.weight" in sd:
                    L["ln_ffm"], L["ffm"] = ln(k + "norm_ff_mem"), ffn(k + "feed_forward_mem")
            self.layers.append(L)
        if self.kind == "oneway_cross_attention":
            # all layers' cross-attention K|V projections side by side: one GEMM hoists them per chain
            self.ca_wkv = bf(th.cat([th.cat([sd[f"{pd}layers.{l}.cross_attn.{n}.0.linear.weight"] for n in ("key", "value")], 0)
                                     for l in range(self.n_layers)], 0))
            self.ca_bkv = f32(th.cat([th.cat([sd[f"{pd}layers.{l}.cross_attn.{n}.0.linear.bias"] for n in ("key", "value")], 0)
                                      for l in range(self.n_layers)], 0))
        self.has_blend = "blend_layer.weight" in sd
        if self.has_blend:
            self.blend_w, self.blend_b = bf(sd["blend_layer.weight"]), f32(sd["blend_layer.bias"])
        self.tmlp = (bf(sd["diffusion_step_encoder.proj.0.weight"]), f32(sd["diffusion_step_encoder.proj.0.bias"]),
                     bf(sd["diffusion_step_encoder.proj.2.weight"]), f32(sd["diffusion_step_encoder.proj.2.bias"]))
        self.pe = positional_table(d, 512).to(device)
        # sinusoidal embedding of the ORIGINAL timestep of every loop index (respace.py:104-113 remap)
        tmap = th.tensor(diffusion.timestep_map, dtype=th.long)
        self.t_embed = step_embedding_table(d, diffusion.original_num_steps)[tmap].to(th.bfloat16).contiguous().to(device)
        self.n_steps = diffusion.num_timesteps


class Op:
    """One kernel launch of the step plan, with the algorithmic work it performs (bench.py's roofline inputs)."""
    __slots__ = ("fn", "kind", "flops", "bytes", "group", "lane")

    def __init__(self, fn, kind, flops=0, nbytes=0):
        self.fn, self.kind, self.flops, self.bytes = fn, kind, flops, nbytes
        # ops of one `group` > 0 form a concurrent region: lane 0 runs on the main stream, lane 1 on the side stream
        self.group, self.lane = 0, 0

    def __call__(self):
        self.fn()


class _Launcher:
    """Thin typed wrappers over the C-ABI; every call goes to libgd_b200.so on the current torch stream."""

    def __init__(self):
        self.lib = gd.load()
        self.keep = []  # tensors that descriptors point into
        self.sm_cap = 0  # > 0: persistent kernels created now are sized for that many SMs (concurrent-lane partitioning)

    @staticmethod
    def stream():
        return th.cuda.current_stream().cuda_stream

    def linear(self, A, W, M, N, K, bias=None, rowbias=None, period=0, offset=0, residual=None, act=gd.ACT_NONE,
               out_f32=None, out_bf16=None, lda=None, ldo32=None, ldo16=None, k_alg=None):
        d = gd.LinearDesc()
        d.A, d.W, d.M, d.N, d.K = _p(A), _p(W), M, N, K
        d.lda, d.ldw = (lda or A.stride(0)), W.stride(0)
        d.bias, d.rowbias, d.rowbias_period, d.rowbias_offset = _p(bias), _p(rowbias), period, offset
        d.residual, d.ldr = _p(residual), (residual.stride(0) if residual is not None else 0)
        d.act = act
        d.out_f32, d.ldo_f32 = _p(out_f32), (ldo32 or (out_f32.stride(0) if out_f32 is not None else 0))
        d.out_bf16, d.ldo_bf16 = _p(out_bf16), (ldo16 or (out_bf16.stride(0) if out_bf16 is not None else 0))
        d.max_ctas = self.sm_cap
        lib = self.lib
        nbytes = 2 * M * K + 2 * N * K + (4 * M * N if residual is not None else 0) + \
            (4 * M * N if out_f32 is not None else 0) + (2 * M * N if out_bf16 is not None else 0)
        return Op(lambda: gd.check(lib.gd_linear_bf16(C.byref(d), self.stream()), "gd_linear_bf16"), "gemm",
                  2 * M * N * (k_alg or K), nbytes)

    def linear_resid_ln(self, A, W, M, N, K, bias, H, ln, xn, ln2=None, split=0):
        """H += A·Wᵀ + bias (in place, fp32) and xn = LayerNorm(H)·gamma + beta (bf16) in one launch; rows >= split use ln2."""
        d = gd.LinearDesc()
        d.A, d.W, d.M, d.N, d.K, d.lda, d.ldw = _p(A), _p(W), M, N, K, A.stride(0), W.stride(0)
        d.bias, d.residual, d.ldr, d.out_f32, d.ldo_f32 = _p(bias), _p(H), H.stride(0), _p(H), H.stride(0)
        l = gd.LnDesc()
        l.gamma, l.beta, l.out_bf16, l.ldo, l.eps = _p(ln[0]), _p(ln[1]), _p(xn), xn.stride(0), 1e-5
        if ln2 is not None:
            l.gamma2, l.beta2, l.split_row = _p(ln2[0]), _p(ln2[1]), split
        lib = self.lib
        nbytes = 2 * M * K + 2 * N * K + 8 * M * N + 2 * M * N
        return Op(lambda: gd.check(lib.gd_linear_resid_ln(C.byref(d), C.byref(l), self.stream()), "gd_linear_resid_ln"),
                  "gemm_ln", 2 * M * N * K + 8 * M * N, nbytes)

    def linear_ln(self, H, ln, W, M, N, K, bias, act=gd.ACT_NONE, out_bf16=None):
        """out = act(LayerNorm(H; ln)·Wᵀ + bias) in one launch (gd_linear_ln_bf16): H fp32 rows, bf16 out."""
        d = gd.LinearDesc()
        d.A, d.W, d.M, d.N, d.K, d.lda, d.ldw = _p(H), _p(W), M, N, K, H.stride(0), W.stride(0)
        d.bias, d.act, d.out_bf16, d.ldo_bf16 = _p(bias), act, _p(out_bf16), out_bf16.stride(0)
        lib, g, b = self.lib, ln[0], ln[1]
        nbytes = 4 * M * K + 2 * N * K + 2 * M * N
        return Op(lambda: gd.check(lib.gd_linear_ln_bf16(C.byref(d), _p(g), _p(b), 1e-5, self.stream()), "gd_linear_ln_bf16"),
                  "gemm", 2 * M * N * K, nbytes)

    def layernorm(self, x, gamma_beta, out, M, D, second=None, split=0):
        """LayerNorm rows; rows >= `split` use the parameter pair `second` when given (one launch for two streams)."""
        lib, g, b = self.lib, gamma_beta[0], gamma_beta[1]
        g2, b2 = (second[0], second[1]) if second is not None else (None, None)
        args = (_p(x), x.stride(0), _p(g), _p(b), _p(g2), _p(b2), split, _p(out), out.stride(0), M, D, 1e-5)
        return Op(lambda: gd.check(lib.gd_layernorm_split(*args, self.stream()), "gd_layernorm"), "layernorm", 8 * M * D, 6 * M * D)

    def attention(self, n_clips, heads, d_k, q, k, v, out, taps, f32in):
        """q/k/v/out: lists of up to two (tensor_view, rows_per_clip) segments; views start at the right column."""
        a = gd.AttnDesc()
        for s, seg in enumerate(q):  # (view, rows[, rows between clips]): a third entry marks a strided (halo) segment
            t, rows = seg[0], seg[1]
            a.q[s], a.q_rows[s], a.q_ld[s] = _p(t), rows, t.stride(0)
            a.q_clip_stride[s] = seg[2] if len(seg) > 2 else 0
        for s, ((tk, rows), (tv, _)) in enumerate(zip(k, v)):
            a.k[s], a.v[s], a.kv_rows[s], a.kv_ld[s] = _p(tk), _p(tv), rows, tk.stride(0)
        for s, (t, _) in enumerate(out):  # fewer outputs than query segments: the rest are conv halos (no output rows)
            a.out[s], a.out_ld[s] = _p(t), t.stride(0)
        a.conv_wq, a.conv_bq, a.conv_wk, a.conv_bk, a.conv_wv, a.conv_bv = [_p(t) for t in taps]
        a.n_clips, a.heads, a.d_k, a.scale = n_clips, heads, d_k, 1.0 / math.sqrt(d_k)
        a.max_ctas_sms = self.sm_cap
        fn = self.lib.gd_dconv_attention_f32in if f32in else self.lib.gd_dconv_attention
        Lq, Lk, dm, es = sum(seg[1] for seg in q[:len(out)]), sum(r for _, r in k), heads * d_k, (4 if f32in else 2)
        flops = n_clips * (4 * Lq * Lk * dm + 6 * dm * (Lq + 2 * Lk))  # QK^T + PV + three 3-tap convs
        nbytes = n_clips * dm * (es * (Lq + 2 * Lk) + 2 * Lq)
        return Op(lambda: gd.check(fn(C.byref(a), self.stream()), "gd_dconv_attention"), "attention", flops, nbytes)


class SamplingChain:
    """One (model, diffusion, batch shape, algorithm) sampling context: buffers + plan + captured graph."""

    def __init__(self, model, diffusion, shape, alg, device, precision="bf16", graph_steps=1, use_graph=True,
                 fuse_ln=False, speech_impl="native", ln_prologue=False, eta=0.0):
        if device.type != "cuda":
            raise gd.GdError("the DDPM sampling path runs only on a CUDA device (sm_100a); there is no CPU fallback")
        # the model is held weakly: the chain cache must not keep a dropped model (and its tape / graph) alive
        self._model_ref = weakref.ref(model)
        self.diffusion, self.alg = diffusion, alg
        self.N, self.C, self.T = shape
        self.device = device
        self.precision = precision
        self.f32act = precision == "fp32act"
        # denoise steps per captured CUDA graph; 0 = the whole chain as ONE graph (the default: 1000 steps = ~190 k kernel nodes,
        # captured once per (model, shape) and replayed for every chain)
        self.graph_steps, self.use_graph = (graph_steps if graph_steps and graph_steps > 0 else diffusion.num_timesteps), use_graph
        if os.environ.get("GD_GRAPH") is not None:  # A/B switch of the determinism harness (profiles/determinism_bisect.py)
            self.use_graph = bool(int(os.environ["GD_GRAPH"]))
        self.L = _Launcher()
        self.W = model.packed_weights(diffusion, device)
        if self.C != self.W.C:
            raise ValueError(f"shape[1]={self.C} does not match the model's d_pose={self.W.C}")
        self.n_steps = diffusion.num_timesteps
        self.eta = float(eta)  # DDIM only (gaussian_diffusion.py:463-467)
        self.tabs = [t.to(device).contiguous() for t in (diffusion.step_tables(alg, self.eta) if alg == "ddim" else
                                                         diffusion.step_tables(alg))]
        self.step = th.zeros(1, dtype=th.int32, device=device)
        # loop index whose step writes the optional outputs (eps / x0 / mean / raw_x0); -1 = every step.  A chain replay
        # needs them for its last step only: 4 x (N,C,T) fp32 stores saved in each of the other 999 steps.
        self.aux_step = th.full((1,), -1, dtype=th.int32, device=device)
        self.x = th.zeros(self.N, self.C, self.T, device=device)
        self.xa = th.zeros(self.N * self.T, _POSE_PAD, device=device, dtype=th.bfloat16)
        self.eps = th.zeros_like(self.x)
        self.x0 = th.zeros_like(self.x)
        self.mean = th.zeros_like(self.x)    # p_mean_variance's `mean` (gaussian_diffusion.py:276-285)
        self.raw_x0 = th.zeros_like(self.x)  # `raw_x_start`: x0 before denoise_fn (:254)
        self._var_tabs = None
        self.tape = None
        self.blend = None
        self.xa_add = None  # Inpaint model: (N,C,T) offset added when the bf16 emb_x operand is built
        self.plan = None
        self.side = th.cuda.Stream(device=device)
        env = os.environ.get  # GD_CONCURRENT / GD_FUSE_LN = 0|1 override the model attributes (A/B measurements)
        self.concurrent = bool(int(env("GD_CONCURRENT", int(getattr(model, "concurrent_streams", True)))))
        # Residual GEMM + following LayerNorm in one kernel (gd_linear_resid_ln).  Off by default: measured on B200 inside
        # the captured chain it loses to the two-kernel form (tedexp 5.44 vs 5.28 ms/step, beat 1.09 vs 1.04) although
        # its kernels sum to less - the stand-alone LayerNorm overlaps with its neighbours, the cluster kernel does not.
        self.fuse_ln = bool(int(env("GD_FUSE_LN", int(fuse_ln))))
        # LayerNorm as the prologue of the consuming GEMM (gd_linear_ln_bf16): no stand-alone LayerNorm launches except the
        # last one (out_layers.0).  bf16 plans only (the fp32-activation parity path keeps fp32 Q|K|V rows).
        self.ln_prologue = bool(int(env("GD_LN_PROLOGUE", int(ln_prologue)))) and not self.f32act \
            and not self.fuse_ln
        # tedexp: hoist LayerNorm + Q|K|V of the layer-0 memory self-attention out of the loop (bf16 plans; GD_HOIST_MEM0=0: off)
        self.hoist_mem0 = bool(int(env("GD_HOIST_MEM0", int(getattr(model, "hoist_mem0", True))))) and not self.f32act \
            and not self.fuse_ln and not self.ln_prologue
        self.encoder_chunk = getattr(model, "encoder_chunk", 16)
        # once-per-clip speech encoder: "native" = ResNetSE-34 on our tensor-core convolutions in split precision (bf16x3:
        # fp32-class products), "native-bf16" = the same with plain bf16 feature maps (faster, ~1e-2 feature error that a
        # random-weight ResNet amplifies), "torch" = the fp32 nn.Module through cuDNN (the fp32-activation parity path)
        self.speech_impl = "torch" if self.f32act else env("GD_SPEECH", speech_impl)
        if self.speech_impl not in ("native", "native-bf16", "torch"):
            raise ValueError(f"speech_impl must be 'native', 'native-bf16' or 'torch', got {self.speech_impl!r}")
        self.native_encoder_chunk = int(env("GD_SPEECH_CHUNK", getattr(model, "native_encoder_chunk", 128)))
        self.graph = None
        self.graph_info = {}
        self.py_denoise = None
        self._plans = {}  # parked plans by key (see begin)
        self._use_tape = True
        self._plan_key = None
        self.Tm = None

    @property
    def model(self):
        m = self._model_ref()
        if m is None:
            raise gd.GdError("the model of this sampling chain has been garbage-collected")
        return m

    # ------------------------------------------------------------------ loop-invariant conditioning
    def _step_token_table(self):
        """z_t for every loop index: Linear -> SiLU -> Linear on the sinusoidal table (nn.py:41-52), two GEMMs."""
        W, d, n = self.W, self.W.d, self.n_steps
        h = th.empty(n, d, device=self.device, dtype=th.bfloat16)
        zt = th.empty(n, d, device=self.device, dtype=th.bfloat16)
        self.L.linear(W.t_embed, W.tmlp[0], n, d, d, bias=W.tmlp[1], act=gd.ACT_SILU, out_bf16=h)()
        self.L.linear(h, W.tmlp[2], n, d, d, bias=W.tmlp[3], out_bf16=zt)()
        return zt

    def _speech_features(self, wav):
        enc = self.model.speech_encoder
        prev = th.backends.cudnn.allow_tf32, th.backends.cuda.matmul.allow_tf32
        th.backends.cudnn.allow_tf32 = th.backends.cuda.matmul.allow_tf32 = False  # reference math is fp32
        try:
            if self.speech_impl != "torch":  # ResNetSE-34 trunk + heads on our kernels (speech_native.py)
                precision = "bf16" if self.speech_impl == "native-bf16" else "bf16x3"
                mel_impl = os.environ.get("GD_MEL", getattr(self.model, "mel_impl", "native"))
                key = (self.model.weights_version, str(self.device), self.native_encoder_chunk, precision, mel_impl)
                cached = getattr(self.model, "_native_speech", None)  # packed once per (weights, device), shared by chains
                if cached is None or cached[0] != key:
                    from .speech_native import NativeSpeechEncoder
                    cached = (key, NativeSpeechEncoder(enc, self.L, self.device, chunk=self.native_encoder_chunk,
                                                       precision=precision, mel_impl=mel_impl))
                    self.model._native_speech = cached
                return cached[1](wav)
            # fixed-size micro-batches (last one zero-padded): cuDNN then runs the same algorithm whatever the batch
            # size, so a clip's conditioning - and with it the whole chain - does not depend on its batch or rank
            chunk = self.encoder_chunk
            outs = []
            with th.no_grad():
                for lo in range(0, wav.shape[0], chunk):
                    w = wav[lo:lo + chunk].float()
                    n = w.shape[0]
                    if n < chunk:
                        w = th.cat([w, w.new_zeros(chunk - n, w.shape[1])], dim=0)
                    outs.append([f[:n] for f in enc(wavform=w)])
            return tuple(th.cat([o[k] for o in outs], dim=0) for k in range(3))
        finally:
            th.backends.cudnn.allow_tf32, th.backends.cuda.matmul.allow_tf32 = prev

    def _conditioning(self, wav):
        W, d, N, dev = self.W, self.W.d, self.N, self.device
        z_low, z_mid, z_high = self._speech_features(wav)
        zt = self._step_token_table()
        n = self.n_steps
        if W.kind == "cross_attention":
            # memory = [z_t ; low ; mid ; high]  (model.py:48-68); positions Tx.. of the joint PE (nn.py:436-442)
            speech = th.cat([z_low, z_mid, z_high], dim=1)  # (N, Ts, d)
            Ts = speech.shape[1]
            Tm = Ts + 1
            sp16 = speech.reshape(N * Ts, d).to(th.bfloat16).contiguous()
            emb = th.empty(N * Ts, d, device=dev)
            self.L.linear(sp16, W.embm_w, N * Ts, d, d, bias=W.embm_b, rowbias=W.pe, period=Ts, offset=self.T + 1,
                          out_f32=emb)()
            mem_init = th.zeros(N, Tm, d, device=dev)
            mem_init[:, 1:] = emb.view(N, Ts, d)
            mem_tab = th.empty(n, d, device=dev)
            self.L.linear(zt, W.embm_w, n, d, d, bias=W.embm_b, rowbias=W.pe, period=1, offset=self.T, out_f32=mem_tab)()
            cond = {"Tm": Tm, "mem_init": mem_init.view(N * Tm, d), "mem_tab": mem_tab}
            if self.hoist_mem0:
                # Layer 0 of the memory stream starts every step from the same rows (only row 0, the timestep token, moves):
                # LayerNorm + fused Q|K|V of the memory self-attention of layer 0 (nn.py:101-103) are evaluated ONCE per chain
                # for the 103 speech rows and once per timestep for row 0 ([n_steps, 3d] table, scattered per step).  Row-wise
                # kernels, so the values are bit-identical to recomputing them in every step.
                ly = W.layers[0]
                xn0 = th.empty(N * Tm, d, device=dev, dtype=th.bfloat16)
                self.L.layernorm(cond["mem_init"], ly["ln_sam"], xn0, N * Tm, d)()
                qkv0 = th.empty(N * Tm, 3 * d, device=dev, dtype=th.bfloat16)
                self.L.linear(xn0, ly["sam"]["wqkv"], N * Tm, 3 * d, d, bias=ly["sam"]["bqkv"], out_bf16=qkv0)()
                xnt = th.empty(n, d, device=dev, dtype=th.bfloat16)
                self.L.layernorm(mem_tab, ly["ln_sam"], xnt, n, d)()
                qkv0_tab = th.empty(n, 3 * d, device=dev, dtype=th.bfloat16)
                self.L.linear(xnt, ly["sam"]["wqkv"], n, 3 * d, d, bias=ly["sam"]["bqkv"], out_bf16=qkv0_tab)()
                cond["qkv0"], cond["qkv0_tab"] = qkv0, qkv0_tab
            return cond
        # oneway: memory = [z_t ; blend(low|mid|high)] -> emb_mem + PE[0..Tm)  (model.py:90-115, nn.py:218-219)
        longest = max(z_low.shape[1], z_mid.shape[1], z_high.shape[1])
        pad = lambda z: th.nn.functional.pad(z, (0, 0, longest - z.shape[1], 0))  # noqa: E731  zero rows in front
        cat16 = th.cat([pad(z_low), pad(z_mid), pad(z_high)], dim=-1).reshape(N * longest, 3 * d).to(th.bfloat16).contiguous()
        z16 = th.empty(N * longest, d, device=dev, dtype=th.bfloat16)
        self.L.linear(cat16, W.blend_w, N * longest, d, 3 * d, bias=W.blend_b, out_bf16=z16)()
        Tm = longest + 1
        mem16 = th.zeros(N, Tm, d, device=dev, dtype=th.bfloat16)
        mem16[:, 1:] = z16.view(N, longest, d)
        memE = th.empty(N * Tm, d, device=dev, dtype=th.bfloat16)  # emb_mem(memory)+PE, bf16 operand of the K|V GEMM
        self.L.linear(mem16.view(N * Tm, d), W.embm_w, N * Tm, d, d, bias=W.embm_b, rowbias=W.pe, period=Tm, offset=0,
                      out_bf16=memE)()
        row0 = th.empty(n, d, device=dev, dtype=th.bfloat16)
        self.L.linear(zt, W.embm_w, n, d, d, bias=W.embm_b, rowbias=W.pe, period=1, offset=0, out_bf16=row0)()
        width = W.n_layers * 2 * d
        kv_dt = th.float32 if self.f32act else th.bfloat16
        kv = th.empty(N * Tm, width, device=dev, dtype=kv_dt)
        kv0 = th.empty(n, width, device=dev, dtype=kv_dt)
        okw = "out_f32" if self.f32act else "out_bf16"
        self.L.linear(memE, W.ca_wkv, N * Tm, width, d, bias=W.ca_bkv, **{okw: kv})()
        self.L.linear(row0, W.ca_wkv, n, width, d, bias=W.ca_bkv, **{okw: kv0})()
        return {"Tm": Tm, "kv": kv, "kv0": kv0}

    # ------------------------------------------------------------------ per-step plan
    def _ddpm_desc(self):
        u = gd.DdpmDesc()
        u.x = _p(self.x)
        u.noise_tape = _p(self.tape) if getattr(self, "_use_tape", True) else None
        u.coef_A, u.coef_B, u.coef_C1, u.coef_C2, u.sigma = [_p(t) for t in self.tabs]
        u.step_ptr = _p(self.step)
        u.n_clips, u.C, u.T = self.N, self.C, self.T
        u.eps_out, u.x0_out = _p(self.eps), _p(self.x0)
        u.mean_out, u.raw_x0_out = _p(self.mean), _p(self.raw_x0)
        u.aux_step_ptr = _p(self.aux_step)
        u.xa_bf16, u.ld_xa = _p(self.xa), _POSE_PAD
        if self.blend is not None:
            u.inpaint_seed, u.inpaint_mask, u.inpaint_factor = _p(self.blend.seed), _p(self.blend.mask), _p(self.blend.factor)
        u.clip_x0 = 0.0
        u.xa_add = _p(self.xa_add)
        return u

    def _attn_block(self, ops, a, rows_lo, rows_hi, segs, xn, qkv, ao, H, n_heads, next_ln=None, ln_in=None):
        """LN'd rows [lo,hi) -> fused QKV GEMM -> dconv attention over `segs` -> out-proj + residual into H
        (+ the LayerNorm that follows, `next_ln`, written to the same rows of xn).  With the LayerNorm-prologue plan the
        QKV GEMM normalises H itself with `ln_in` and no LayerNorm follows."""
        d, L = self.W.d, self.L
        M = rows_hi - rows_lo
        okw = "out_f32" if self.f32act else "out_bf16"
        if self.ln_prologue:
            ops.append(L.linear_ln(H[rows_lo:rows_hi], ln_in, a["wqkv"], M, 3 * d, d, a["bqkv"], out_bf16=qkv[rows_lo:rows_hi]))
            next_ln = None
        else:
            ops.append(L.linear(xn[rows_lo:rows_hi], a["wqkv"], M, 3 * d, d, bias=a["bqkv"], **{okw: qkv[rows_lo:rows_hi]}))
        q = [(qkv[lo:, 0:], r) for lo, r in segs]
        k = [(qkv[lo:, d:], r) for lo, r in segs]
        v = [(qkv[lo:, 2 * d:], r) for lo, r in segs]
        o = [(ao[lo:], r) for lo, r in segs]
        ops.append(L.attention(self.N, n_heads, d // n_heads, q, k, v, o, a["taps"], self.f32act))
        self._resid(ops, ao[rows_lo:rows_hi], a["wo"], a["bo"], M, d, H[rows_lo:rows_hi], next_ln, xn[rows_lo:rows_hi])

    def _resid(self, ops, A, W, bias, M, K, H, ln, xn, ln2=None, split=0):
        """Residual GEMM H += A·Wᵀ + b, fused with the following LayerNorm when `ln` is given (and fusion is on)."""
        d, L = self.W.d, self.L
        if ln is not None and self.fuse_ln:
            ops.append(L.linear_resid_ln(A, W, M, d, K, bias, H, ln, xn, ln2=ln2, split=split))
            return
        ops.append(L.linear(A, W, M, d, K, bias=bias, residual=H, out_f32=H))
        if ln is not None:
            if ln2 is None or split >= M:
                ops.append(L.layernorm(H, ln, xn, M, d))
            else:  # pose rows and memory rows in ONE launch, each with its own gamma / beta
                ops.append(L.layernorm(H, ln, xn, M, d, second=ln2, split=split))

    def _ffn_block(self, ops, f, lo, hi, xn, hid, H, next_ln=None, ln_in=None):
        """xn rows [lo,hi) already hold LN(H): up-projection + ReLU², down-projection + residual (+ the next LayerNorm).
        LayerNorm-prologue plan: the up-projection normalises H rows with `ln_in` itself."""
        d, L = self.W.d, self.L
        M = hi - lo
        if self.ln_prologue:
            ops.append(L.linear_ln(H[lo:hi], ln_in, f["w1"], M, 4 * d, d, f["b1"], act=gd.ACT_RELU2, out_bf16=hid[lo:hi]))
            next_ln = None
        else:
            ops.append(L.linear(xn[lo:hi], f["w1"], M, 4 * d, d, bias=f["b1"], act=gd.ACT_RELU2, out_bf16=hid[lo:hi]))
        self._resid(ops, hid[lo:hi], f["w2"], f["b2"], M, 4 * d, H[lo:hi], next_ln, xn[lo:hi])

    def _build_plan(self, cond):
        W, L, N, T, d, dev = self.W, self.L, self.N, self.T, self.W.d, self.device
        heads, lib = W.heads, self.L.lib
        Tm = cond["Tm"]
        Mx = N * T
        ops = []
        qkv_dt = th.float32 if self.f32act else th.bfloat16
        if W.kind == "cross_attention":
            Mm = N * Tm
            R = Mx + Mm
            H = th.empty(R, d, device=dev)
            xn = th.empty(R, d, device=dev, dtype=th.bfloat16)
            qkv = th.empty(R, 3 * d, device=dev, dtype=qkv_dt)
            ao = th.empty(R, d, device=dev, dtype=th.bfloat16)
            hid = th.empty(R, 4 * d, device=dev, dtype=th.bfloat16)
            X, Mem = H[:Mx], H[Mx:]
            # The pose stream and the memory stream only meet in the joint attention, so [FFN(l-1), self-attn(l)] of the
            # two streams are tagged as concurrent regions: small pose-stream kernels fill the tails of the memory-stream ones.
            def tag(start, group, lane):
                for op in ops[start:]:
                    op.group, op.lane = group, lane
            region = 1
            # SM partitioning of the concurrent regions (GD_SM_PARTITION / model.sm_partition): the persistent kernels of the
            # pose lane are sized for Mx/R of the SMs and those of the memory lane for the rest, so the two lanes run side by
            # side instead of queueing behind each other's 148-CTA grids
            part = bool(int(os.environ.get("GD_SM_PARTITION", int(getattr(self.model, "sm_partition", SM_PARTITION_DEFAULT))))) \
                and self.concurrent
            n_sm = th.cuda.get_device_properties(dev).multi_processor_count
            cap0 = max(8, int(round(n_sm * Mx / R))) if part else 0
            cap1 = (n_sm - cap0) if part else 0
            a_sc = (_p(Mem), _p(cond["mem_init"]), _p(cond["mem_tab"]), _p(self.step), N, Tm, 0, d, d)
            ops.append(Op(lambda: gd.check(lib.gd_scatter_step_row_f32(*a_sc, L.stream()), "gd_scatter_step_row_f32"),
                          "scatter", 0, 8 * Mm * d))
            tag(len(ops) - 1, region, 1)
            L.sm_cap = cap0
            ops.append(L.linear(self.xa, W.embx_w, Mx, d, _POSE_PAD, bias=W.embx_b, rowbias=W.pe, period=T, offset=0,
                                out_f32=X, k_alg=self.C))
            L.sm_cap = 0
            tag(len(ops) - 1, region, 0)
            # Default plan: every LayerNorm is a launch right behind the residual GEMM that completes its input (the first one
            # of each stream here).  LayerNorm-prologue plan: the consuming GEMM normalises H itself (`ln_in`).
            if not self.ln_prologue:
                ops.append(L.layernorm(X, W.layers[0]["ln_sa"], xn[:Mx], Mx, d))
                tag(len(ops) - 1, region, 0)
                if "qkv0" not in cond:
                    ops.append(L.layernorm(Mem, W.layers[0]["ln_sam"], xn[Mx:], Mm, d))
                    tag(len(ops) - 1, region, 1)
            for li, ly in enumerate(W.layers):
                last = li == W.n_layers - 1
                nxt = None if last else W.layers[li + 1]
                s0 = len(ops)
                L.sm_cap = cap0
                self._attn_block(ops, ly["sa"], 0, Mx, [(0, T)], xn, qkv, ao, H, heads, next_ln=ly["ln_ca"], ln_in=ly["ln_sa"])
                tag(s0, region, 0)
                s0 = len(ops)
                L.sm_cap = cap1
                if li == 0 and "qkv0" in cond:
                    # layer 0: Q|K|V of the memory rows come from the per-chain buffer; only row 0 of every clip is per step
                    q0, w3 = cond["qkv0"], 3 * d
                    a_q0 = (_p(q0), _p(cond["qkv0_tab"]), _p(self.step), N, Tm, 0, w3, w3)
                    ops.append(Op(lambda: gd.check(lib.gd_scatter_step_row_bf16(*a_q0, L.stream()), "gd_scatter_step_row_bf16"),
                                  "scatter", 0, 4 * N * w3))
                    a0 = ly["sam"]
                    ops.append(L.attention(N, heads, d // heads, [(q0[:, 0:], Tm)], [(q0[:, d:], Tm)], [(q0[:, 2 * d:], Tm)],
                                           [(ao[Mx:], Tm)], a0["taps"], False))
                    self._resid(ops, ao[Mx:R], a0["wo"], a0["bo"], Mm, d, H[Mx:R], ly["ln_ca"], xn[Mx:R])
                else:
                    self._attn_block(ops, ly["sam"], Mx, R, [(Mx, Tm)], xn, qkv, ao, H, heads, next_ln=ly["ln_ca"], ln_in=ly["ln_sam"])
                tag(s0, region, 1)
                L.sm_cap = 0
                region += 1
                # joint attention over [x ; memory] (nn.py:105-113); last layer only the pose rows are read afterwards
                a = ly["ca"]
                okw = "out_f32" if self.f32act else "out_bf16"
                if self.ln_prologue:
                    ops.append(L.linear_ln(H, ly["ln_ca"], a["wqkv"], R, 3 * d, d, a["bqkv"], out_bf16=qkv))
                else:
                    ops.append(L.linear(xn, a["wqkv"], R, 3 * d, d, bias=a["bqkv"], **{okw: qkv}))
                segs = [(0, T), (Mx, Tm)]
                qs = [(qkv[lo:, 0:], r) for lo, r in segs]
                if last:
                    # only the pose rows are read afterwards (nn.py:445-447), but the symmetric conv3 of the last pose
                    # frame reaches the first memory row (nn.py:105-113, transformer.py:19-44): memory row 0 of every
                    # clip rides along as a one-row halo segment without output
                    qs = [qs[0], (qkv[Mx:, 0:], 1, Tm)]
                os_ = [(ao[lo:], r) for lo, r in (segs[:1] if last else segs)]
                ks = [(qkv[lo:, d:], r) for lo, r in segs]
                vs = [(qkv[lo:, 2 * d:], r) for lo, r in segs]
                ops.append(L.attention(N, heads, d // heads, qs, ks, vs, os_, a["taps"], self.f32act))
                Ro = Mx if last else R
                self._resid(ops, ao[:Ro], a["wo"], a["bo"], Ro, d, H[:Ro], None if self.ln_prologue else ly["ln_ff"], xn[:Ro],
                            ln2=ly.get("ln_ffm"), split=Mx)
                s0 = len(ops)
                L.sm_cap = cap0 if "ffm" in ly else 0
                self._ffn_block(ops, ly["ff"], 0, Mx, xn, hid, H, next_ln=(W.out_ln if last else nxt["ln_sa"]), ln_in=ly["ln_ff"])
                if "ffm" in ly:
                    tag(s0, region, 0)
                    s0 = len(ops)
                    L.sm_cap = cap1
                    self._ffn_block(ops, ly["ffm"], Mx, R, xn, hid, H, next_ln=(None if last else nxt["ln_sam"]), ln_in=ly["ln_ffm"])
                    tag(s0, region, 1)
                L.sm_cap = 0
        else:
            X = th.empty(Mx, d, device=dev)
            H = X
            xn = th.empty(Mx, d, device=dev, dtype=th.bfloat16)
            qkv = th.empty(Mx, 3 * d, device=dev, dtype=qkv_dt)
            ao = th.empty(Mx, d, device=dev, dtype=th.bfloat16)
            hid = th.empty(Mx, 4 * d, device=dev, dtype=th.bfloat16)
            kv, kv0 = cond["kv"], cond["kv0"]
            width = kv.shape[1]
            if self.f32act:
                a_sc = (_p(kv), None, _p(kv0), _p(self.step), N, Tm, 0, width, width)
                ops.append(Op(lambda: gd.check(lib.gd_scatter_step_row_f32(*a_sc, L.stream()), "gd_scatter_step_row_f32"),
                              "scatter", 0, 8 * N * width))
            else:
                a_sc = (_p(kv), _p(kv0), _p(self.step), N, Tm, 0, width, width)
                ops.append(Op(lambda: gd.check(lib.gd_scatter_step_row_bf16(*a_sc, L.stream()), "gd_scatter_step_row_bf16"),
                              "scatter", 0, 4 * N * width))
            ops.append(L.linear(self.xa, W.embx_w, Mx, d, _POSE_PAD, bias=W.embx_b, rowbias=W.pe, period=T, offset=0,
                                out_f32=X, k_alg=self.C))
            okw = "out_f32" if self.f32act else "out_bf16"
            if not self.ln_prologue:
                ops.append(L.layernorm(X, W.layers[0]["ln_sa"], xn, Mx, d))
            for li, ly in enumerate(W.layers):
                last = li == W.n_layers - 1
                self._attn_block(ops, ly["sa"], 0, Mx, [(0, T)], xn, qkv, ao, H, heads, next_ln=ly["ln_ca"], ln_in=ly["ln_sa"])
                # cross attention: queries from the pose rows, K|V hoisted per chain (memory is never updated, nn.py:160-162)
                a = ly["ca"]
                if self.ln_prologue:
                    ops.append(L.linear_ln(X, ly["ln_ca"], a["wqkv"][:d], Mx, d, d, a["bqkv"][:d], out_bf16=qkv[:, :d]))
                else:
                    ops.append(L.linear(xn, a["wqkv"][:d], Mx, d, d, bias=a["bqkv"][:d], **{okw: qkv[:, :d]}))
                kcol = li * 2 * d
                ops.append(L.attention(N, heads, d // heads, [(qkv[:, 0:], T)], [(kv[:, kcol:], Tm)], [(kv[:, kcol + d:], Tm)],
                                       [(ao, T)], a["taps"], self.f32act))
                self._resid(ops, ao, a["wo"], a["bo"], Mx, d, X, None if self.ln_prologue else ly["ln_ff"], xn)
                self._ffn_block(ops, ly["ff"], 0, Mx, xn, hid, H, next_ln=(W.out_ln if last else W.layers[li + 1]["ln_sa"]),
                                ln_in=ly["ln_ff"])
        if self.ln_prologue:  # the one LayerNorm left: out_layers.0 in front of the fused projection + DDPM update
            ops.append(L.layernorm(H[:Mx], W.out_ln, xn[:Mx], Mx, d))
        # xn[:Mx] holds out_layers.0 LayerNorm(x) (written by the last down-projection)
        dd = gd.LinearDesc()
        dd.A, dd.W, dd.M, dd.N, dd.K, dd.lda, dd.ldw, dd.bias = _p(xn), _p(W.out_w), Mx, _POSE_PAD, d, d, d, _p(W.out_b)
        self._ddpm = self._ddpm_desc()
        elems = N * self.C * T
        ops.append(Op(lambda: gd.check(lib.gd_linear_ddpm(C.byref(dd), C.byref(self._ddpm), L.stream()), "gd_linear_ddpm"),
                      "gemm_ddpm", 2 * Mx * d * self.C, 2 * Mx * d + 2 * _POSE_PAD * d + 4 * elems * 5 + 2 * Mx * _POSE_PAD))
        a_st = (_p(self.step), -1)
        ops.append(Op(lambda: gd.check(lib.gd_step_add(*a_st, L.stream()), "gd_step_add"), "step", 0, 8))
        self._buffers = (H, xn, qkv, ao, hid, cond)
        return ops

    # ------------------------------------------------------------------ public driver
    def begin(self, x_T, wav, denoise_fn=None, noise_tape=None, need_tape=True, input_offset=None, rng_consumed=False):
        """Load x_T, compute the conditioning once, (re)build the plan, reset the step counter.  `input_offset`
        (N,C,T) is the Inpaint model's loop-invariant offset of the denoiser input (model.py:161-165).  `rng_consumed`:
        the caller (SplitChain) has already drawn this chain's share of the reference's per-step random numbers."""
        from .diffusion import InpaintBlend
        # Any other callable (gaussian_diffusion.py:256-257 accepts arbitrary `denoise_fn`) cannot run inside a captured
        # graph: the chain then steps eagerly - the denoiser still on the kernels, eps through a plain final projection, the
        # update as the reference's own elementwise torch ops around the Python call (step_eager_callable).
        self.py_denoise = None
        if denoise_fn is not None and not isinstance(denoise_fn, InpaintBlend):
            if not callable(denoise_fn):
                raise TypeError("denoise_fn must be callable")
            self.py_denoise, denoise_fn = denoise_fn, None
        if tuple(x_T.shape) != (self.N, self.C, self.T):
            raise ValueError(f"noise shape {tuple(x_T.shape)} != {(self.N, self.C, self.T)}")
        assert wav.dim() == 2 and wav.shape[0] == self.N, f"Wav dim should be (N,T). Got: {tuple(wav.shape)}"
        dev = self.device
        tape_shape = (self.n_steps, self.N, self.C, self.T)
        if need_tape:
            if noise_tape is None:
                # the reference draws one randn_like per step in loop order (gaussian_diffusion.py:326); same calls,
                # same generator stream.  Loop position k uses index i = n-1-k.
                if self.tape is None or self.tape.shape != tape_shape:
                    self.tape = th.empty(tape_shape, device=dev)
                for k in range(self.n_steps):  # in place: same generator stream as randn_like, no staging copy
                    self.tape[self.n_steps - 1 - k].normal_()
            else:
                assert tuple(noise_tape.shape) == tape_shape, f"noise_tape must be {tape_shape}"
                tp = noise_tape.to(dev).float().flip(0).contiguous()  # loop order -> index order
                if self.tape is not None and self.tape.shape == tape_shape:
                    self.tape.copy_(tp)
                else:
                    self.tape = tp
        elif self.alg == "ddim" and not rng_consumed and getattr(self.model, "match_reference_rng", True):
            # the reference's ddim_sample draws th.randn_like(x) at every step even though eta = 0 multiplies it away
            # (gaussian_diffusion.py:475): consume the same draws so that a seeded sequence of calls (generate_sequence
            # windows, repeated generate_sample) sees the same x_T stream afterwards
            scratch = th.empty(self.N, self.C, self.T, device=dev)
            for _ in range(self.n_steps):
                scratch.normal_()
        blend_key = None if denoise_fn is None else "blend"
        cond = self._conditioning(wav)
        if input_offset is not None:
            if tuple(input_offset.shape) != (self.N, self.C, self.T):
                raise ValueError(f"input_offset shape {tuple(input_offset.shape)} != {(self.N, self.C, self.T)}")
            if self.xa_add is None:
                self.xa_add = th.empty(self.N, self.C, self.T, device=dev)
            self.xa_add.copy_(input_offset.float())
        elif self.xa_add is not None:
            self.xa_add = None
        # A plan (launch list + descriptors + activation buffers + captured graph) is specific to the memory length, to whether
        # the in-paint blend / the noise tape / the input offset are wired into the update epilogue, and to the tape and offset
        # buffers it points at.  Plans are kept side by side: generate_sequence goes window 0 (no blend) -> window 1
        # (blend) -> ..., eval code alternates sampling and teacher-forced steps; each plan is built and captured once.
        key = (cond["Tm"], blend_key, need_tape, _p(self.tape) if need_tape else 0, _p(self.xa_add))
        if self._plan_key is not None and key != self._plan_key:  # park the active plan
            self._plans[self._plan_key] = (self.plan, self.graph, self.graph_info, self._buffers, self._ddpm, self.blend, self.Tm)
        st = self._plans.pop(key, None) if key != self._plan_key else None
        if self.plan is None or (key != self._plan_key and st is None):
            self.blend = denoise_fn
            self.Tm = cond["Tm"]
            self._use_tape = need_tape
            self.plan = self._build_plan(cond)
            self._plan_key, self.graph, self.graph_info = key, None, {}
        else:
            if st is not None:
                self.plan, self.graph, self.graph_info, self._buffers, self._ddpm, self.blend, self.Tm = st
                self._plan_key = key
            # same plan/graph: refresh the buffers the captured kernels read
            old = self._buffers[5]
            for name, val in cond.items():
                if isinstance(val, th.Tensor):
                    old[name].copy_(val)
            if denoise_fn is not None:
                self.blend.seed.copy_(denoise_fn.seed)
                self.blend.mask.copy_(denoise_fn.mask)
                self.blend.factor.copy_(denoise_fn.factor)
        self._eps_rows = None  # (the eager Python-denoise_fn projection is bound to the active plan's buffers)
        while len(self._plans) > 3:  # bounded: the oldest parked plan goes
            self._plans.pop(next(iter(self._plans)))
        self.x.copy_(x_T.float())
        self._pack_pose_rows()
        self.step.fill_(self.n_steps - 1)
        self._pos = 0  # steps executed since begin()

    def step_eager(self):
        """Enqueue one denoise step.  Concurrent regions fork onto the side stream and join back (inside a capture this
        becomes graph-level parallelism)."""
        main = th.cuda.current_stream()
        cur = 0
        for op in self.plan:
            if op.group != cur:
                if cur > 0:
                    main.wait_stream(self.side)
                if op.group > 0 and self.concurrent:
                    self.side.wait_stream(main)
                cur = op.group if self.concurrent else 0
            if cur > 0 and op.lane == 1:
                with th.cuda.stream(self.side):
                    op()
            else:
                op()
        if cur > 0:
            main.wait_stream(self.side)

    def step_eager_callable(self):
        """One denoise step with an arbitrary Python `denoise_fn` between pred_x_start and the posterior mean
        (p_mean_variance, gaussian_diffusion.py:252-259): every kernel of the plan except the fused projection+update, a plain
        final projection to eps, then the update as torch elementwise ops in the reference's order."""
        W, d, Mx = self.W, self.W.d, self.N * self.T
        if getattr(self, "_eps_rows", None) is None or self._eps_rows.shape[0] != Mx:
            self._eps_rows = th.empty(Mx, _POSE_PAD, device=self.device)
            xn = self._buffers[1]
            self._proj_op = self.L.linear(xn[:Mx], W.out_w, Mx, _POSE_PAD, d, bias=W.out_b, out_f32=self._eps_rows)
        for op in self.plan[:-2]:  # (the concurrent lanes run back to back here)
            op()
        self._proj_op()
        i = int(self.step.item())
        A, B, C1, C2, sig = (t[i] for t in self.tabs)
        eps = self._eps_rows.view(self.N, self.T, _POSE_PAD)[:, :, :self.C].transpose(1, 2).contiguous()
        x = self.x
        raw = A * x - B * eps
        x0 = self.py_denoise(raw.clone())
        mean = C1 * x0 + C2 * x
        z = self.tape[i] if (self.tape is not None and i != 0) else None
        new_x = mean + sig * z if z is not None else mean + 0.0
        self.eps.copy_(eps); self.raw_x0.copy_(raw); self.x0.copy_(x0); self.mean.copy_(mean)
        self.x.copy_(new_x)
        self._pack_pose_rows()
        self.step.fill_(i - 1)

    def set_state(self, x, i):
        """Teacher forcing: overwrite the sample and the loop index (parity harness)."""
        self.x.copy_(x.to(self.device).float())
        self._pack_pose_rows()
        self.step.fill_(int(i))
        self._pos = self.n_steps - 1 - int(i)

    def _pack_pose_rows(self):
        gd.check(self.L.lib.gd_pack_pose_rows_add(_p(self.x), _p(self.xa_add), _p(self.xa), self.N, self.C, self.T, _POSE_PAD,
                                                  self.L.stream()), "gd_pack_pose_rows_add")

    def _warm_step(self):
        """One eager step on a side stream warms every kernel (func attributes, tensor-map encode path); state restored."""
        saved = (self.x.clone(), self.xa.clone(), self.step.clone())
        s = th.cuda.Stream(device=self.device)
        s.wait_stream(th.cuda.current_stream())
        with th.cuda.stream(s):
            self.step_eager()
        th.cuda.current_stream().wait_stream(s)
        th.cuda.synchronize()
        self.x.copy_(saved[0]); self.xa.copy_(saved[1]); self.step.copy_(saved[2])
        return saved

    def _ensure_graph(self):
        if self.graph is not None or not self.use_graph:
            return
        saved = self._warm_step()
        import time
        free0 = th.cuda.mem_get_info(self.device)[0]
        t0 = time.perf_counter()
        g = th.cuda.CUDAGraph()
        with th.cuda.graph(g):
            for _ in range(self.graph_steps):
                self.step_eager()
        th.cuda.synchronize()
        t1 = time.perf_counter()
        self.x.copy_(saved[0]); self.xa.copy_(saved[1]); self.step.copy_(saved[2])
        self.graph = g
        g.replay()  # the first replay uploads the executable graph to the device: count it as part of the one-off cost
        th.cuda.synchronize()
        self.x.copy_(saved[0]); self.xa.copy_(saved[1]); self.step.copy_(saved[2])
        # first-call cost of the public API: capture + instantiate + upload, and what the graph holds
        self.graph_info = {"steps_per_graph": self.graph_steps, "kernel_nodes": self.graph_steps * len(self.plan),
                           "capture_s": round(t1 - t0, 3), "first_replay_s": round(time.perf_counter() - t1, 3),
                           "device_bytes": int(max(free0 - th.cuda.mem_get_info(self.device)[0], 0))}

    def _result(self, i=0):
        """The dict of the step that used loop index `i` (the last executed step), keys as the reference's:
        p_sample -> sample, mean, variance, log_variance, eps, pred_x_start, raw_x_start (gaussian_diffusion.py:276-285,329);
        ddim_sample -> sample, pred_x_start (:484)."""
        if self._var_tabs is None:  # float64 table -> fp32 at the gather, then broadcast (`_extract_into_tensor`, :681-694)
            d = self.diffusion
            self._var_tabs = (th.from_numpy(d.posterior_variance).to(self.device).float(),
                              th.from_numpy(d.posterior_log_variance_clipped).to(self.device).float())
        var, logvar = (t[i] + th.zeros_like(self.x) for t in self._var_tabs)
        out = {"sample": self.x, "variance": var, "log_variance": logvar, "eps": self.eps, "pred_x_start": self.x0,
               "raw_x_start": self.raw_x0}
        if self.alg == "ddpm":  # the DDIM tables fold the posterior mean away (eta = 0: sample == mean_pred, :476-484)
            out["mean"] = self.mean
        return out

    def run(self, progress=False, n_steps=None):
        """Run the remaining chain (or `n_steps` steps). Returns the last step's dict."""
        total = self.n_steps if n_steps is None else n_steps
        if self.py_denoise is not None:
            for _ in range(total):
                self.step_eager_callable()
            self._pos += total
            return self._result(max(self.n_steps - self._pos, 0))
        self.aux_step.fill_(max(self.n_steps - self._pos - total, 0))  # only the last step of this run hands back its dict
        if self.use_graph:
            self._ensure_graph()
            done = 0
            while total - done >= self.graph_steps:
                self.graph.replay()
                done += self.graph_steps
            for _ in range(total - done):
                self.step_eager()
        else:
            for _ in range(total):
                self.step_eager()
        self._pos += total
        self.aux_step.fill_(-1)
        return self._result(max(self.n_steps - self._pos, 0))

    def iterate(self, progress=False):
        """Progressive form (p_sample_loop_progressive): yields cloned per-step dicts."""
        for i in range(self.n_steps - 1, -1, -1):
            if self.py_denoise is not None:
                self.step_eager_callable()
            else:
                self.step_eager()
            self._pos += 1
            yield {k: v.clone() for k, v in self._result(i).items()}


class SplitChain:
    """K independent sub-chains over contiguous clip slices, captured as PARALLEL BRANCHES of one CUDA graph.

    Clips never interact, so a batch may be sampled as several smaller batches.  At small batches a
    denoise step is bound by the serial critical path of its 49-192 kernels, not by throughput (B200, beat-ours:
    64 clips 0.21 ms/step, 128 clips 0.30 ms/step - most SMs idle): two or four sub-chains running side by side fill the
    idle SMs, and because every kernel is batch-invariant the poses are bit-identical to the un-split run
    (tests/test_batch_parity_gpu.py).  Only the whole-chain entry points (`p_sample_loop` / `ddim_sample_loop`) use it; the
    progressive / teacher-forced / bpd paths keep a single chain."""

    def __init__(self, model, diffusion, shape, alg, device, parts, **opts):
        self.N, self.C, self.T = shape
        base, extra = divmod(self.N, parts)
        self.bounds, lo = [], 0
        for k in range(parts):
            hi = lo + base + (1 if k < extra else 0)
            self.bounds.append((lo, hi))
            lo = hi
        opts = dict(opts, use_graph=False)  # the children never capture on their own
        self.children = [SamplingChain(model, diffusion, (hi - lo, self.C, self.T), alg, device, **opts) for lo, hi in self.bounds]
        self.streams = [th.cuda.Stream(device=device) for _ in self.children]
        self.device, self.alg, self.diffusion = device, alg, diffusion
        self.n_steps = diffusion.num_timesteps
        self.speech_impl = self.children[0].speech_impl
        self.graph, self.graph_info, self._keys = None, {}, None
        self.parts = parts

    @property
    def plan(self):  # all launches of one denoise step (bench.py counts FLOPs / kernels from it)
        return [op for ch in self.children for op in ch.plan]

    def begin(self, x_T, wav, denoise_fn=None, noise_tape=None, need_tape=True, input_offset=None):
        dev, n = self.device, self.n_steps
        if denoise_fn is not None and not hasattr(denoise_fn, "slice"):
            raise NotImplementedError("a Python denoise_fn needs the single (eager) chain: set model.sub_chains = 1")
        if need_tape and noise_tape is None:
            # the same draws, in the same order, as the single chain (one (N,C,T) normal_() per step, loop order)
            noise_tape = th.empty(n, self.N, self.C, self.T, device=dev)
            for k in range(n):
                noise_tape[k].normal_()
        elif self.alg == "ddim" and getattr(self.children[0].model, "match_reference_rng", True):
            scratch = th.empty(self.N, self.C, self.T, device=dev)
            for _ in range(n):
                scratch.normal_()
        for ch, (lo, hi) in zip(self.children, self.bounds):  # the draws above were made once for the whole batch
            ch.begin(x_T[lo:hi], wav[lo:hi], denoise_fn=None if denoise_fn is None else denoise_fn.slice(lo, hi),
                     noise_tape=None if noise_tape is None else noise_tape[:, lo:hi], need_tape=need_tape,
                     input_offset=None if input_offset is None else input_offset[lo:hi], rng_consumed=True)
        keys = tuple(ch._plan_key for ch in self.children)
        if keys != self._keys:
            self._keys, self.graph = keys, None

    def _capture(self):
        import time
        saved = [ch._warm_step() for ch in self.children]
        free0 = th.cuda.mem_get_info(self.device)[0]
        t0 = time.perf_counter()
        g = th.cuda.CUDAGraph()
        with th.cuda.graph(g):
            main = th.cuda.current_stream()
            for ch, s in zip(self.children, self.streams):
                s.wait_stream(main)
                with th.cuda.stream(s):
                    for _ in range(self.n_steps):
                        ch.step_eager()
            for s in self.streams:
                main.wait_stream(s)
        th.cuda.synchronize()
        t1 = time.perf_counter()
        self.graph = g

        def restore():
            for ch, sv in zip(self.children, saved):
                ch.x.copy_(sv[0]); ch.xa.copy_(sv[1]); ch.step.copy_(sv[2])
        restore()
        g.replay()
        th.cuda.synchronize()
        restore()
        self.graph_info = {"steps_per_graph": self.n_steps, "parallel_sub_chains": self.parts,
                           "kernel_nodes": self.n_steps * len(self.plan), "capture_s": round(t1 - t0, 3),
                           "first_replay_s": round(time.perf_counter() - t1, 3),
                           "device_bytes": int(max(free0 - th.cuda.mem_get_info(self.device)[0], 0))}

    def run(self, progress=False, n_steps=None):
        if n_steps is not None and n_steps != self.n_steps:
            raise ValueError("a split chain runs whole chains only")
        if self.graph is None:
            self._capture()
        for ch in self.children:
            ch.aux_step.fill_(0)  # only the last step hands back its dict
        self.graph.replay()
        outs = []
        for ch in self.children:
            ch._pos = self.n_steps
            ch.aux_step.fill_(-1)
            outs.append(ch._result(0))
        return {k: th.cat([o[k] for o in outs], dim=0) for k in outs[0]}


def sub_chain_count(model, shape, alg):
    """How many parallel sub-chains a whole-chain sampling call is split into (1 = none).  `model.sub_chains` / GD_SUBCHAINS:
    an int, or "auto" (default): split while a sub-batch keeps at least AUTO_MIN_ROWS token rows and the batch is small
    enough that one chain leaves SMs idle (measured on B200)."""
    want = os.environ.get("GD_SUBCHAINS", getattr(model, "sub_chains", "auto"))
    N, _, T = shape
    if str(want) != "auto":
        return max(1, min(int(want), N))
    rows = N * T
    parts = 1
    while parts < AUTO_MAX_PARTS and rows // (2 * parts) >= AUTO_MIN_ROWS and rows <= AUTO_SPLIT_BELOW_ROWS:
        parts *= 2
    return parts


AUTO_MIN_ROWS, AUTO_SPLIT_BELOW_ROWS, AUTO_MAX_PARTS = 1 << 30, 0, 4  # auto splitting off until measured (set below)


_CHAINS = {}
# LayerNorm-prologue plan as the default of new chains (set after the A/B measurement on B200)
LN_PROLOGUE_DEFAULT = False
# SM partitioning of the concurrent pose / memory lanes of the tedexp plan (set after the A/B measurement on B200)
SM_PARTITION_DEFAULT = False


def release_chains(model=None):
    """Drop the cached sampling contexts of `model` (all models if None): buffers, noise tapes and captured graphs."""
    for k in [k for k in _CHAINS if model is None or k[0] == id(model)]:
        del _CHAINS[k]
    import gc
    gc.collect()
    if th.cuda.is_available():
        th.cuda.empty_cache()


def chain_for(model, diffusion, shape, alg, device, allow_split=False, **kw):
    """Cache of sampling contexts keyed by (model, diffusion, shape, algorithm): buffers and graphs are reused.  Entries of
    a model are dropped when the model is garbage-collected, when its weights change (load_state_dict / .to()) and when
    another batch shape is requested, so the cache holds at most one batch shape per live model."""
    device = th.device(device)
    if device.type == "cuda" and device.index is None:
        device = th.device("cuda", th.cuda.current_device())
    opts = dict(precision=getattr(model, "precision", "bf16"), graph_steps=getattr(model, "graph_steps", 0),
                use_graph=getattr(model, "use_graph", True), fuse_ln=getattr(model, "fuse_layernorm", False),
                speech_impl=getattr(model, "speech_impl", "native"), ln_prologue=getattr(model, "ln_prologue", LN_PROLOGUE_DEFAULT),
                eta=0.0)
    opts.update(kw)
    parts = sub_chain_count(model, shape, alg) if (allow_split and opts["use_graph"] and not opts["graph_steps"]) else 1
    key = (id(model), id(diffusion), shape, alg, str(device), model.weights_version, tuple(sorted(opts.items())), parts)
    ch = _CHAINS.get(key)
    if ch is None:
        for k in [k for k in _CHAINS if k[0] == id(model) and (k[2] != shape or k[5] != model.weights_version)]:
            del _CHAINS[k]  # keep one batch shape per model resident; chains of superseded weights are dead
        ch = (SamplingChain(model, diffusion, shape, alg, device, **opts) if parts <= 1 else
              SplitChain(model, diffusion, shape, alg, device, parts, **opts))
        _CHAINS[key] = ch
        if not getattr(model, "_gd_chain_finalizer", False):
            weakref.finalize(model, release_chains_by_id, id(model))
            model._gd_chain_finalizer = True
    return ch


def release_chains_by_id(model_id):
    for k in [k for k in _CHAINS if k[0] == model_id]:
        del _CHAINS[k]
