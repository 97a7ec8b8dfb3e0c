"""Plain-torch statements of the speech-encoder entry points of include/gd_b200.h (test infrastructure).

Two uses: (1) on the GPU the real kernels are compared with these on the same descriptors; (2) on CPU `FakeLib` stands in
for libgd_b200.so so that the host logic of gesture_b200.speech_native (weight packing, grid geometry, tap shifts, head
merge) is checked end to end against the fp32 `SpeechEncoder` module without a GPU.  bf16 storage is emulated by rounding.
"""
import torch as th


def bf16_round(t):
    return t.to(th.bfloat16).float()


def join(t, c, split):
    """value of a feature-map row tensor [rows, c] or [rows, hi(c) | lo(c)] as fp32 [rows, c]."""
    return t[:, :c].float() + t[:, c:2 * c].float() if split else t[:, :c].float()


def put(out, index, v, c, split):
    """store fp32 values [*, c] into rows `index` of a plain / split feature map (bf16 emulated by the tensor dtype
    or, for fp32 stand-ins on CPU, by rounding)."""
    hi = v.to(th.bfloat16)
    out[index, :c] = hi.to(out.dtype)
    if split:
        out[index, c:2 * c] = (v - hi.float()).to(th.bfloat16).to(out.dtype)


def conv_taps_ref(inp, W, n_images, grid_h, grid_w, taps, k_per_tap, bias, scale, shift, relu, window, stride, out, out_strides,
                  c_store, split_out, walk=0):
    """gd_conv_taps_bf16 on tensors: inp [rows, in_ld], W [c_out, n_taps*k_per_tap], out [*, out_ld] modified in place."""
    rows, in_ld = n_images * grid_h * grid_w, inp.shape[1]
    A = inp[:rows].float()
    Wf = W.float()
    if walk == 2:   # per 64-channel block the W columns are [Whi(64) | Wlo(64)]; input columns hi, lo, hi
        c = in_ld // 2
        blk = th.arange(c // 64, device=inp.device)[:, None] * 64 + th.arange(64, device=inp.device)[None]   # [blocks, 64]
        cols = th.cat([blk, blk + c, blk], dim=1).reshape(-1)                  # input column per product term
        wbase = (th.arange(c // 64, device=inp.device)[:, None] * 128 + th.arange(64, device=inp.device)[None])
        wcols = th.cat([wbase, wbase, wbase + 64], dim=1).reshape(-1)
    else:           # generic (and walk 1, which is the same arithmetic): k-th element reads column k mod in_ld
        cols = th.arange(k_per_tap, device=inp.device) % in_ld
        wcols = th.arange(k_per_tap, device=inp.device)
    acc = th.zeros(rows, W.shape[0], dtype=th.float32, device=inp.device)
    for t, sh in enumerate(taps):
        shifted = th.zeros_like(A)
        lo, hi = max(0, -sh), min(rows, rows - sh)
        if hi > lo:
            shifted[lo:hi] = A[lo + sh:hi + sh]
        acc += shifted[:, cols] @ Wf[:, t * k_per_tap + wcols].T
    if bias is not None:
        acc = acc + bias
    if relu:
        acc = acc.clamp_min(0)
    acc = acc * scale + shift
    r = th.arange(rows, device=inp.device)
    img, rem = r // (grid_h * grid_w), r % (grid_h * grid_w)
    y, x = rem // grid_w, rem % grid_w
    y0, y1, x0, x1 = window
    keep = (y >= y0) & (y <= y1) & (x >= x0) & (x <= x1) & ((y - y0) % stride == 0) & ((x - x0) % stride == 0)
    si, sy, sx, off = out_strides
    orow = img * si + ((y - y0) // stride) * sy + ((x - x0) // stride) * sx + off
    put(out, orow[keep], acc[keep][:, :c_store], c_store, split_out)
    return out


def stem_ref(mel, w, bias, scale, shift, out, c_pad, split=0):
    """gd_speech_stem: mel (n, H, W) fp32 -> bordered channel-last rows [n*(H+2)*(W+2), c_pad (x2 if split)]."""
    n, H, W = mel.shape
    c = w.shape[0]
    y = th.nn.functional.conv2d(mel[:, None], w.view(c, 1, 3, 3), bias, padding=1).clamp_min(0)
    y = y * scale[None, :, None, None] + shift[None, :, None, None]
    v = th.zeros(n, H, W, c_pad, device=mel.device)
    v[..., :c] = y.permute(0, 2, 3, 1)
    idx = th.arange(n * (H + 2) * (W + 2), device=mel.device).view(n, H + 2, W + 2)[:, 1:-1, 1:-1].reshape(-1)
    put(out, idx, v.reshape(-1, c_pad), c_pad, split)
    return out


def se_gate_ref(y, n_images, grid_h, grid_w, c, split, c_real, w1, b1, w2, b2, gate):
    g = join(y[:n_images * grid_h * grid_w], c, split).view(n_images, grid_h * grid_w, c)
    mean = g.sum(dim=1)[:, :c_real] / ((grid_h - 2) * (grid_w - 2))
    a = th.sigmoid(th.relu(mean @ w1.T + b1) @ w2.T + b2)
    gate[:n_images] = 0
    gate[:n_images, :c_real] = a
    return gate


def se_residual_relu_ref(y, res, gate, out, n_images, grid_h, grid_w, c, split=0):
    rows = n_images * grid_h * grid_w
    v = lambda t: join(t[:rows], c, split).view(n_images, grid_h, grid_w, c)  # noqa: E731
    o = th.relu(gate[:n_images, None, None, :c] * v(y) + v(res))
    idx = th.arange(rows, device=y.device).view(n_images, grid_h, grid_w)[:, 1:-1, 1:-1].reshape(-1)
    put(out, idx, o[:, 1:-1, 1:-1].reshape(-1, c), c, split)
    return out


def pixel_shuffle_ref(inp, out, n_images, H, W, c_in, r, c_out_pad, split=0):
    o = out[:n_images * H * r * W * r]
    o.zero_()
    for pl in range(2 if split else 1):
        plane = inp[:n_images * (H + 2) * (W + 2), pl * c_in:(pl + 1) * c_in]
        feat = plane.reshape(n_images, H + 2, W + 2, c_in)[:, 1:-1, 1:-1].permute(0, 3, 1, 2)
        sh = th.nn.functional.pixel_shuffle(feat.float(), r)  # (n, c_in/r^2, H*r, W*r)
        o[:, pl * c_out_pad:pl * c_out_pad + sh.shape[1]] = sh.permute(0, 2, 3, 1).reshape(-1, sh.shape[1]).to(out.dtype)
    return out


def mel_power_ref(wav, window, fb, preemph, add_eps):
    """gd_mel_power: pre-emphasis (reflect pad 1) -> |STFT|^2 (n_fft 1024, hop 512, centre reflect) -> mel filterbank + eps."""
    x = th.nn.functional.pad(wav[:, None], (1, 0), "reflect")[:, 0]
    y = x[:, 1:] - preemph * x[:, :-1]
    spec = th.stft(y, 1024, hop_length=512, win_length=1024, window=window, center=True, pad_mode="reflect",
                   normalized=False, onesided=True, return_complex=True).abs().pow(2.0)
    return th.matmul(spec.transpose(-1, -2), fb).transpose(-1, -2) + add_eps


def instance_norm_ref(x, eps):
    mean = x.mean(dim=-1, keepdim=True)
    var = x.var(dim=-1, unbiased=False, keepdim=True)
    return (x - mean) / th.sqrt(var + eps)


class FakeLauncher:
    """Stands in for engine._Launcher + libgd_b200.so on CPU: every 'kernel' is the torch statement above.  Tensors are
    found through the pointers the host code puts into the descriptors (registered by `track`)."""

    def __init__(self):
        self.lib = self
        self.tensors = {}
        self.calls = []

    def track(self, t):
        if t is not None:
            self.tensors[t.data_ptr()] = t
        return None if t is None else t.data_ptr()

    def _t(self, ptr):
        return None if not ptr else self.tensors[ptr]

    @staticmethod
    def stream():
        return None

    def linear(self, A, W, M, N, K, bias=None, out_f32=None, **kw):
        assert not kw and A.shape[1] == K and W.shape == (N, K)

        def run():
            self.calls.append("gd_linear_bf16")
            out_f32[:M] = A[:M].float() @ W.float().T + bias
        return run

    def gd_conv_taps_bf16(self, ref, stream):
        d = ref._obj
        self.calls.append("gd_conv_taps_bf16")
        inp, W, out = self._t(d.inp), self._t(d.W), self._t(d.out)
        assert inp.shape[1] == d.in_ld and W.shape == (d.c_out, d.n_taps * d.k_per_tap)
        assert d.in_ld % 64 == 0 and d.k_per_tap % 64 == 0 and d.c_out % 64 == 0 and d.c_store % 32 == 0
        out2 = out.view(-1, d.out_ld)
        conv_taps_ref(inp, W, d.n_images, d.grid_h, d.grid_w, [d.tap_shift[i] for i in range(d.n_taps)], d.k_per_tap,
                      self._t(d.bias), self._t(d.scale), self._t(d.shift), d.relu, (d.y0, d.y1, d.x0, d.x1), d.stride, out2,
                      (d.out_img_stride, d.out_y_stride, d.out_x_stride, d.out_offset), d.c_store, d.split_out, d.walk)
        return 0

    def gd_speech_stem(self, mel, w, b, sc, sh, out, n, H, W, c_real, c_pad, split, stream):
        self.calls.append("gd_speech_stem")
        stem_ref(self._t(mel)[:n], self._t(w), self._t(b), self._t(sc), self._t(sh), self._t(out), c_pad, split)
        return 0

    @staticmethod
    def gd_se_gate_scratch_bytes(n, gh, gw, c):
        return 4 * (65536 + n * ((gh * gw + 255) // 256) * c)

    def gd_se_gate(self, y, n, gh, gw, c, split, c_real, c_hidden, w1, b1, w2, b2, gate, scratch, scratch_bytes, stream):
        self.calls.append("gd_se_gate")
        assert scratch_bytes >= self.gd_se_gate_scratch_bytes(n, gh, gw, c)
        assert self._t(w1).shape == (c_hidden, c_real)
        se_gate_ref(self._t(y), n, gh, gw, c, split, c_real, self._t(w1), self._t(b1), self._t(w2), self._t(b2), self._t(gate))
        return 0

    def gd_se_residual_relu(self, y, res, gate, out, n, gh, gw, c, split, stream):
        self.calls.append("gd_se_residual_relu")
        se_residual_relu_ref(self._t(y), self._t(res), self._t(gate), self._t(out), n, gh, gw, c, split)
        return 0

    def gd_pixel_shuffle_rows(self, inp, out, n, H, W, c_in, r, c_out_pad, split, stream):
        self.calls.append("gd_pixel_shuffle_rows")
        pixel_shuffle_ref(self._t(inp), self._t(out), n, H, W, c_in, r, c_out_pad, split)
        return 0

    def gd_mel_power(self, wav, n, length, window, twiddle, fb, fb_range, preemph, add_eps, mel, stream):
        self.calls.append("gd_mel_power")
        fbt, rg = self._t(fb), self._t(fb_range)
        for m in range(128):  # the band table must cover every non-zero weight
            lo, hi = int(rg[m, 0]), int(rg[m, 1])
            assert fbt[:lo, m].abs().sum() == 0 and fbt[hi + 1:, m].abs().sum() == 0
        tw = self._t(twiddle).double()
        k = th.arange(512, dtype=th.float64) * (2 * th.pi / 1024)
        assert (tw[:, 0] - th.cos(k)).abs().max() < 1e-7 and (tw[:, 1] + th.sin(k)).abs().max() < 1e-7
        self._t(mel).copy_(mel_power_ref(self._t(wav)[:n, :length], self._t(window), fbt, preemph, add_eps))
        return 0

    def gd_instance_norm_rows(self, x, rows, length, eps, stream):
        self.calls.append("gd_instance_norm_rows")
        t = self._t(x)
        t.copy_(instance_norm_ref(t.reshape(rows, length), eps).reshape(t.shape))
        return 0

    def gd_last_error(self):
        return b""
