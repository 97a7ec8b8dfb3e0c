"""Denoiser module: the drop-in for the reference's `Speech2GestureModel` / `Speech2GestureModelV2`
(`models/model.py:6-117`).  It is an `nn.Module` with the reference's parameter names (`load_state_dict` of a
reference checkpoint works unchanged, `main.py:110-115`), but its forward is the sm_100a kernel chain, not torch ops.
"""
import torch as th
import torch.nn as nn

from .modules import PoseDecoderParams, SpeechEncoder, StepEncoderParams


class Speech2GestureDenoiser(nn.Module):
    """model(x_t (N,C,T) fp32, t (N,) int64, wav=(N,T_wav)) -> eps (N,C,T) fp32   (models/model.py:12-15).

    model_type 'default' = tedexp-ours wrapper (memory = [z_t ; low ; mid ; high], model.py:41-73);
    's2g_v2' = beat-ours wrapper (adds `blend_layer`, memory = [z_t ; blend(low|mid|high)], model.py:76-117);
    'inpaint' = the default wrapper whose input is offset by `proj([inpaint_pose*mask | mask])`, a zero-initialised
    3-layer SiLU MLP (model.py:120-166); it takes the extra kwargs inpaint_pose (T,N,C) and inpaint_mask (T,N,1).
    Attributes `precision` ('bf16' | 'fp32act'), `graph_steps`, `use_graph` steer the engine.
    """

    def __init__(self, model_type, d_pose, d_model, speech_encoder, pose_decoder, diffusion_step_encoder, dropout_prob=0.0,
                 pose_seed_len=None):
        super().__init__()
        if model_type not in ("default", "s2g_v2", "inpaint"):
            raise ValueError(f"Unsupported model_type {model_type}")
        self.model_type = model_type
        self.diffusion_step_encoder = diffusion_step_encoder
        self.speech_encoder = speech_encoder
        self.pose_decoder = pose_decoder
        self.d_pose, self.d_model = d_pose, d_model
        if model_type == "s2g_v2":
            self.blend_layer = nn.Linear(3 * d_model, d_model)
        if model_type == "inpaint":
            self.pose_seed_len = pose_seed_len
            self.proj = nn.Sequential(nn.Linear(d_pose + 1, d_model), nn.SiLU(), nn.Linear(d_model, d_model), nn.SiLU(),
                                      nn.Linear(d_model, d_pose), nn.Dropout(dropout_prob))
            for m in self.proj:  # zero init as GLIDE (model.py:147-153)
                if isinstance(m, nn.Linear):
                    m.weight.data.zero_()
                    m.bias.data.zero_()
        self.precision, self.graph_steps, self.use_graph = "bf16", 0, True  # graph_steps 0: whole chain in one CUDA graph
        self.weights_version = 0
        self._packed = None
        self._diffusion = None
        self.register_load_state_dict_post_hook(lambda module, incompatible: module.invalidate_packed())

    # -- weights ----------------------------------------------------------------------------------
    def invalidate_packed(self):
        """Call after changing parameters in place; load_state_dict / .to() do it automatically."""
        self.weights_version += 1
        self._packed = None

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self.invalidate_packed()
        return out

    def packed_weights(self, diffusion, device):
        from .engine import PackedWeights
        key = (self.weights_version, id(diffusion), str(device))
        if self._packed is None or self._packed[0] != key:
            self._packed = (key, PackedWeights(self, diffusion, device))
        return self._packed[1]

    def bind_diffusion(self, diffusion):
        """The timestep-token table depends on the diffusion's timestep map; create_model binds it."""
        self._diffusion = diffusion
        return self

    def count_learnable_parameters(self):
        n = sum(p.numel() for p in self.parameters() if p.requires_grad)
        print("[Info] Number of parameters: {:,}".format(n))
        return n

    def input_offset(self, model_kwargs):
        """Inpaint model: proj([inpaint_pose*mask | mask]) as an (N,C,T) fp32 tensor (model.py:161-165).  It does not
        depend on the timestep, so it is evaluated once per chain (three small fp32 Linears) and the kernels add it to
        the sample when they build the bf16 operand of emb_x.  None for the other model types."""
        if self.model_type != "inpaint":
            return None
        pose, mask = model_kwargs.get("inpaint_pose"), model_kwargs.get("inpaint_mask")
        if pose is None or mask is None:
            raise ValueError("the inpaint model needs model_kwargs['inpaint_pose'] (T,N,C) and ['inpaint_mask'] (T,N,1)")
        dev = next(self.parameters()).device
        pose, mask = pose.to(dev).float(), mask.to(dev).float()
        prev = th.backends.cuda.matmul.allow_tf32
        th.backends.cuda.matmul.allow_tf32 = False  # reference math is fp32
        try:
            with th.no_grad():
                delta = self.proj(th.cat([pose * mask, mask], dim=-1))  # (T,N,C)
        finally:
            th.backends.cuda.matmul.allow_tf32 = prev
        return delta.permute(1, 2, 0).contiguous()  # -> (N,C,T)

    # -- forward: one denoiser evaluation -------------------------------------------------------------
    @th.no_grad()
    def forward(self, x_t, t, **model_kwargs):
        """Single eps prediction through the same kernels the chain uses (all clips must share one t, which is
        how every sampling loop of the reference calls it, gaussian_diffusion.py:402).  `t` is the ORIGINAL
        timestep (what `_WrappedModel` passes, respace.py:110-113)."""
        from .engine import chain_for
        if self._diffusion is None:
            raise RuntimeError("model is not bound to a diffusion process; build it with create_model(...)")
        if x_t.device.type != "cuda":
            from ._lib import GdError
            raise GdError("the denoiser runs only on a CUDA device (sm_100a kernels); there is no CPU fallback")
        t0 = int(t.reshape(-1)[0].item())
        if not bool((t == t0).all()):
            # per-clip timesteps (models/model.py:12-15: `t` is (N,); training and calc_bpd's callers pass mixed values):
            # clips sharing a timestep go through the kernels together; the kernels are batch-invariant, so every clip's
            # eps is what a uniform-t call would give it
            out = th.empty_like(x_t, dtype=th.float32)
            for tv in th.unique(t).tolist():
                idx = (t == tv).nonzero(as_tuple=True)[0]
                kw = {k: (v[idx] if k == "wav" else v[:, idx]) for k, v in model_kwargs.items()}  # inpaint_* are (T,N,.)
                out[idx] = self.forward(x_t[idx].contiguous(), t[idx], **kw)
            return out
        i = self._diffusion.timestep_map.index(t0)
        chain = chain_for(self, self._diffusion, tuple(x_t.shape), "ddpm", x_t.device, use_graph=False)
        chain.begin(x_t, model_kwargs["wav"].to(x_t.device), need_tape=False, input_offset=self.input_offset(model_kwargs))
        chain.set_state(x_t, i)
        chain.step_eager()
        return chain.eps.clone()
