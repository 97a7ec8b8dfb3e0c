"""GPU parity of the once-per-clip speech encoder: each kernel of include/gd_b200.h's speech section against its plain
torch statement (tests/speech_ref.py), then the whole native encoder against the fp32 `SpeechEncoder` module."""
import ctypes as C
import math

import pytest
import torch as th

import speech_ref as ref
from test_speech_host_cpu import randomise_batchnorm

pytestmark = pytest.mark.gpu


def _stream():
    return th.cuda.current_stream().cuda_stream


def _conv(gd, inp, W, n, gh, gw, taps, bias, scale, shift, relu, window, stride, out, out_strides, split=0, c_store=None, walk=0):
    d = gd.ConvDesc()
    d.inp, d.W, d.n_images, d.grid_h, d.grid_w = inp.data_ptr(), W.data_ptr(), n, gh, gw
    d.in_ld, d.k_per_tap, d.c_out, d.n_taps = inp.shape[1], W.shape[1] // len(taps), W.shape[0], len(taps)
    d.c_store, d.split_out, d.walk = c_store or W.shape[0], split, walk
    for i, t in enumerate(taps):
        d.tap_shift[i] = t
    d.bias, d.scale, d.shift, d.relu = gd.ptr(bias), scale.data_ptr(), shift.data_ptr(), int(relu)
    d.y0, d.y1, d.x0, d.x1 = window
    d.stride = stride
    d.out, d.out_ld = out.data_ptr(), out.shape[1]
    d.out_img_stride, d.out_y_stride, d.out_x_stride, d.out_offset = out_strides
    gd.check(gd.load().gd_conv_taps_bf16(C.byref(d), _stream()), "gd_conv_taps_bf16")
    th.cuda.synchronize()


def _bordered(n, H, W, c, g, c_real=None, split=0):
    """random feature map on a zero-bordered grid, channels >= c_real zero; bf16 [rows, c] or split [rows, hi(c) | lo(c)]."""
    x = th.zeros(n, H + 2, W + 2, c, device="cuda")
    x[:, 1:-1, 1:-1, :c_real or c] = th.randn(n, H, W, c_real or c, device="cuda", generator=g)
    x = x.reshape(-1, c)
    if not split:
        return x.bfloat16()
    hi = x.bfloat16()
    return th.cat([hi, (x - hi.float()).bfloat16()], dim=1).contiguous()


@pytest.mark.parametrize("n,H,W,ci,co,k,stride,relu,with_bias", [
    (3, 16, 9, 64, 64, 3, 1, True, False),      # layer1-like block convolution
    (2, 16, 9, 64, 128, 3, 2, True, False),     # first convolution of a stage (stride 2)
    (2, 16, 9, 64, 128, 1, 2, False, False),    # its 1x1 down-sample branch
    (5, 12, 7, 128, 256, 3, 1, False, True),    # wide output tile, bias
    (70, 32, 16, 128, 128, 3, 1, False, False),  # enough m-tiles for the 128-wide tile path (>= 148)
    (40, 16, 8, 256, 256, 3, 1, True, False),   # long K (2304), 256-wide tiles
])
def test_conv_taps_same_padding(gd, n, H, W, ci, co, k, stride, relu, with_bias):
    g = th.Generator(device="cuda").manual_seed(n * 100 + ci + co + k)
    x = _bordered(n, H, W, ci, g)
    w4 = th.randn(co, ci, k, k, device="cuda", generator=g) / math.sqrt(ci * k * k)
    Wp = w4.permute(0, 2, 3, 1).reshape(co, k * k * ci).bfloat16().contiguous()
    kpt = ci
    bias = th.randn(co, device="cuda", generator=g) if with_bias else None
    scale = th.rand(co, device="cuda", generator=g) + 0.5
    shift = th.randn(co, device="cuda", generator=g) * 0.1
    Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
    gw = W + 2
    taps = [0] if k == 1 else [(ky - 1) * gw + (kx - 1) for ky in range(3) for kx in range(3)]
    geo = ((Ho + 2) * (Wo + 2), Wo + 2, 1, Wo + 3)
    out = th.full((n * (Ho + 2) * (Wo + 2), co), 7.0, device="cuda", dtype=th.bfloat16)
    _conv(gd, x, Wp, n, H + 2, gw, taps, bias, scale, shift, relu, (1, H, 1, W), stride, out, geo)
    want = th.full_like(out, 7.0)
    ref.conv_taps_ref(x, Wp, n, H + 2, gw, taps, kpt, bias, scale, shift, relu, (1, H, 1, W), stride, want, geo, co, 0)
    # independent statement: F.conv2d on the same bf16-rounded operands
    img = x.float().view(n, H + 2, W + 2, ci)[:, 1:-1, 1:-1].permute(0, 3, 1, 2)
    y = th.nn.functional.conv2d(img, Wp.float().view(co, k, k, ci).permute(0, 3, 1, 2), bias, stride=stride, padding=k // 2)
    y = (y.clamp_min(0) if relu else y) * scale[None, :, None, None] + shift[None, :, None, None]
    got = out.view(n, Ho + 2, Wo + 2, co)
    assert (got[:, 1:-1, 1:-1].float() - y.permute(0, 2, 3, 1)).abs().max().item() < 3e-2
    assert (out.float() - want.float()).abs().max().item() < 3e-2
    border = th.ones(Ho + 2, Wo + 2, dtype=th.bool, device="cuda")
    border[1:-1, 1:-1] = False
    assert (got[:, border] == 7.0).all(), "border pixels must never be written"


@pytest.mark.parametrize("n,H,W,ci,co,k,stride,relu", [
    (3, 16, 9, 32, 32, 3, 1, True),      # layer1: one 128-byte row holds both planes, two k-blocks per tap
    (2, 16, 9, 32, 64, 3, 2, True),      # 32 -> 64 channels with stride 2
    (2, 16, 9, 32, 64, 1, 2, False),     # 1x1 down-sample branch
    (70, 32, 16, 128, 128, 3, 1, False),
    (40, 16, 8, 256, 256, 3, 1, True),   # K = 9 * 768
])
@pytest.mark.parametrize("reuse", [True, False])
def test_conv_taps_split_precision(gd, n, H, W, ci, co, k, stride, relu, reuse):
    """bf16x3: split feature maps and split weights reproduce the fp32 convolution to ~1e-5 relative."""
    from gesture_b200.speech_native import _pack_conv, _pad_to
    g = th.Generator(device="cuda").manual_seed(n * 100 + ci + co + k + 1)
    x = _bordered(n, H, W, ci, g, split=1)
    w4 = th.randn(co, ci, k, k, device="cuda", generator=g) / math.sqrt(ci * k * k)
    co_pad = _pad_to(co)
    Wp, kpt, walk = _pack_conv(w4, ci, co_pad, 1, reuse)
    assert walk == (0 if not reuse else 1 if ci == 32 else 2)
    pad = lambda v: th.cat([v, v.new_zeros(co_pad - co)])  # noqa: E731
    scale, shift = pad(th.rand(co, device="cuda", generator=g) + 0.5), pad(th.randn(co, device="cuda", generator=g) * 0.1)
    Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
    gw = W + 2
    taps = [0] if k == 1 else [(ky - 1) * gw + (kx - 1) for ky in range(3) for kx in range(3)]
    geo = ((Ho + 2) * (Wo + 2), Wo + 2, 1, Wo + 3)
    out = th.zeros(n * (Ho + 2) * (Wo + 2), 2 * co, device="cuda", dtype=th.bfloat16)
    _conv(gd, x, Wp, n, H + 2, gw, taps, None, scale, shift, relu, (1, H, 1, W), stride, out, geo, split=1, c_store=co, walk=walk)
    xv = ref.join(x, ci, 1).view(n, H + 2, W + 2, ci)[:, 1:-1, 1:-1].permute(0, 3, 1, 2).double()
    y = th.nn.functional.conv2d(xv, w4.double(), None, stride=stride, padding=k // 2)
    y = (y.clamp_min(0) if relu else y) * scale[None, :co, None, None] + shift[None, :co, None, None]
    got = ref.join(out, co, 1).view(n, Ho + 2, Wo + 2, co)
    err = ((got[:, 1:-1, 1:-1].double() - y.permute(0, 2, 3, 1)).norm() / y.norm()).item()
    assert err < 3e-5, err
    assert got[:, 0].abs().max().item() == 0 and got[:, :, 0].abs().max().item() == 0


def test_conv_taps_valid_head_layout(gd):
    """3x3 valid convolution on an unbordered grid, output transposed to [image][x][y][c] (the pyramid heads)."""
    n, gh, gw, c = 3, 20, 12, 64
    g = th.Generator(device="cuda").manual_seed(3)
    x = th.randn(n * gh * gw, c, device="cuda", generator=g).bfloat16()
    w4 = th.randn(c, c, 3, 3, device="cuda", generator=g) / math.sqrt(9 * c)
    Wp = w4.permute(0, 2, 3, 1).reshape(c, 9 * c).bfloat16().contiguous()
    bias, scale, shift = (th.randn(c, device="cuda", generator=g) for _ in range(3))
    T, hh = gw - 2, gh - 2
    out = th.zeros(n * T * hh, c, device="cuda", dtype=th.bfloat16)
    taps = [(ky - 1) * gw + (kx - 1) for ky in range(3) for kx in range(3)]
    _conv(gd, x, Wp, n, gh, gw, taps, bias, scale, shift, True, (1, gh - 2, 1, gw - 2), 1, out, (T * hh, 1, hh, 0))
    img = x.float().view(n, gh, gw, c).permute(0, 3, 1, 2)
    y = th.nn.functional.conv2d(img, Wp.float().view(c, 3, 3, c).permute(0, 3, 1, 2), bias).clamp_min(0)
    y = y * scale[None, :, None, None] + shift[None, :, None, None]          # (n, c, hh, T)
    assert (out.view(n, T, hh, c).float() - y.permute(0, 3, 2, 1)).abs().max().item() < 4e-2


def test_conv_taps_rejects_bad_arguments(gd):
    x = th.zeros(100, 64, device="cuda", dtype=th.bfloat16)
    w = th.zeros(64, 64, device="cuda", dtype=th.bfloat16)
    v = th.zeros(64, device="cuda")
    with pytest.raises(gd.GdError):
        _conv(gd, x[:, :32].contiguous(), w[:, :32].contiguous(), 1, 10, 10, [0], None, v, v, False, (1, 8, 1, 8), 1, x, (100, 10, 1, 11))
    with pytest.raises(gd.GdError):
        _conv(gd, x, w, 1, 10, 10, [0], None, v, v, False, (1, 10, 1, 8), 1, x, (100, 10, 1, 11))  # window outside grid
    with pytest.raises(gd.GdError):
        _conv(gd, x, w, 1, 10, 10, [0], None, v, v, False, (1, 8, 1, 8), 3, x, (100, 10, 1, 11))   # stride 3
    with pytest.raises(gd.GdError):
        _conv(gd, x, w, 1, 10, 10, [0], None, v, v, False, (1, 8, 1, 8), 1, x, (100, 10, 1, 11), c_store=48)
    with pytest.raises(gd.GdError):
        _conv(gd, x, w, 1, 10, 10, [0], None, v, v, False, (1, 8, 1, 8), 1, x, (100, 10, 1, 11), walk=2)  # in_ld 64 is not 2c


@pytest.mark.parametrize("split", [0, 1])
def test_stem_gate_tail_shuffle(gd, split):
    lib = gd.load()
    g = th.Generator(device="cuda").manual_seed(11)
    n, H, W, c_real, c = 3, 128, 13, 32, (32 if split else 64)
    planes = 2 if split else 1
    tol = 1e-5 if split else 3e-2
    mel = th.randn(n, H, W, device="cuda", generator=g)
    w, b = th.randn(c_real, 9, device="cuda", generator=g) / 3, th.randn(c_real, device="cuda", generator=g)
    sc, sh = th.rand(c_real, device="cuda", generator=g) + 0.5, th.randn(c_real, device="cuda", generator=g)
    out = th.zeros(n * (H + 2) * (W + 2), planes * c, device="cuda", dtype=th.bfloat16)
    gd.check(lib.gd_speech_stem(mel.data_ptr(), w.data_ptr(), b.data_ptr(), sc.data_ptr(), sh.data_ptr(), out.data_ptr(),
                                n, H, W, c_real, c, split, _stream()), "gd_speech_stem")
    want = ref.stem_ref(mel, w, b, sc, sh, th.zeros_like(out), c, split)
    assert (ref.join(out, c, split) - ref.join(want, c, split)).abs().max().item() < tol * 10

    # squeeze-excite gate on that map (32 real channels, hidden 4) and on a 256-channel map
    for (y, gh, gw, cc, cr) in [(out, H + 2, W + 2, c, c_real), (_bordered(4, 16, 8, 256, g, split=split), 18, 10, 256, 256)]:
        ch, ni = cr // 8, y.shape[0] // (gh * gw)
        w1, b1 = th.randn(ch, cr, device="cuda", generator=g) / math.sqrt(cr), th.randn(ch, device="cuda", generator=g)
        w2, b2 = th.randn(cr, ch, device="cuda", generator=g) / math.sqrt(ch), th.randn(cr, device="cuda", generator=g)
        gate = th.full((ni, cc), -1.0, device="cuda")
        scratch = th.zeros(lib.gd_se_gate_scratch_bytes(ni, gh, gw, cc) // 4, device="cuda")
        for _ in range(2):  # twice: the kernel must leave its counters ready for the next launch
            gd.check(lib.gd_se_gate(y.data_ptr(), ni, gh, gw, cc, split, cr, ch, w1.data_ptr(), b1.data_ptr(), w2.data_ptr(),
                                    b2.data_ptr(), gate.data_ptr(), scratch.data_ptr(), scratch.numel() * 4, _stream()), "gd_se_gate")
        with pytest.raises(gd.GdError):
            lib_rc = lib.gd_se_gate(y.data_ptr(), ni, gh, gw, cc, split, cr, ch, w1.data_ptr(), b1.data_ptr(), w2.data_ptr(),
                                    b2.data_ptr(), gate.data_ptr(), scratch.data_ptr(), 16, _stream())
            gd.check(lib_rc, "gd_se_gate")
        want_gate = ref.se_gate_ref(y, ni, gh, gw, cc, split, cr, w1, b1, w2, b2, th.zeros_like(gate))
        assert (gate - want_gate).abs().max().item() < 1e-4
        # block tail
        res = _bordered(ni, gh - 2, gw - 2, cc, g, cr, split=split)
        o = th.full_like(y, 5.0)
        gd.check(lib.gd_se_residual_relu(y.data_ptr(), res.data_ptr(), gate.data_ptr(), o.data_ptr(), ni, gh, gw, cc, split,
                                         _stream()), "gd_se_residual_relu")
        want_o = ref.se_residual_relu_ref(y, res, want_gate, th.full_like(y, 5.0), ni, gh, gw, cc, split)
        assert (ref.join(o, cc, split) - ref.join(want_o, cc, split)).abs().max().item() < tol * 10
        ov = o.view(ni, gh, gw, planes * cc)
        assert (ov[:, 0] == 5.0).all() and (ov[:, :, -1] == 5.0).all()

    # pixel shuffles of the mid / high heads
    for (Hs, Ws, ci, r) in [(32, 5, 128, 2), (16, 3, 256, 4)]:
        x = _bordered(2, Hs, Ws, ci, g, split=split)
        co = 32 if split else 64
        o = th.full((2 * Hs * r * Ws * r, planes * co), 9.0, device="cuda", dtype=th.bfloat16)
        gd.check(lib.gd_pixel_shuffle_rows(x.data_ptr(), o.data_ptr(), 2, Hs, Ws, ci, r, co, split, _stream()), "gd_pixel_shuffle_rows")
        want_s = ref.pixel_shuffle_ref(x, th.zeros_like(o), 2, Hs, Ws, ci, r, co, split)
        assert th.equal(o, want_s)


# default-init weights with random BatchNorm statistics; bf16x3 is limited by the bf16 merged head weights (~2e-3)
ENC_TOL = {"bf16x3": 4e-3, "bf16": 2e-2}


@pytest.mark.parametrize("precision", ["bf16x3", "bf16"])
@pytest.mark.parametrize("d_model,n,wav_len,chunk", [(256, 5, 32000, 4), (512, 3, 36266, 64), (256, 2, 128000, 64)])
def test_native_encoder_matches_fp32_module(gd, d_model, n, wav_len, chunk, precision):
    from gesture_b200.engine import _Launcher
    from gesture_b200.modules import SpeechEncoder
    from gesture_b200.speech_native import NativeSpeechEncoder
    th.manual_seed(0)
    enc = SpeechEncoder(d_model).eval()
    randomise_batchnorm(enc)
    enc.cuda()
    wav = th.randn(n, wav_len, device="cuda", generator=th.Generator(device="cuda").manual_seed(4))
    prev = th.backends.cudnn.allow_tf32, th.backends.cuda.matmul.allow_tf32
    th.backends.cudnn.allow_tf32 = th.backends.cuda.matmul.allow_tf32 = False
    try:
        with th.no_grad():
            want = enc(wavform=wav)
        native = NativeSpeechEncoder(enc, _Launcher(), th.device("cuda", 0), chunk=chunk, precision=precision)
        got = native(wav)
        again = native(wav[n - 1:])          # another batch composition: bit-identical per clip
    finally:
        th.backends.cudnn.allow_tf32, th.backends.cuda.matmul.allow_tf32 = prev
    for a, b, c in zip(want, got, again):
        assert a.shape == b.shape
        rel = ((a - b).norm() / a.norm()).item()
        assert rel < ENC_TOL[precision], (precision, rel)
        assert th.equal(b[n - 1:], c)


@pytest.mark.parametrize("n,wav_len", [(3, 32000), (2, 36266), (1, 5000), (2, 128000)])
def test_mel_front_end(gd, n, wav_len):
    """gd_mel_power + gd_instance_norm_rows vs torch.stft / matmul / InstanceNorm1d, on noise and on a signal whose high
    bands are 80 dB below the low ones (the per-bin normalisation amplifies errors in weak bands)."""
    from gesture_b200.engine import _Launcher
    from gesture_b200.modules import SpeechEncoder
    from gesture_b200.speech_native import NativeSpeechEncoder
    th.manual_seed(0)
    enc = SpeechEncoder(256).eval().cuda()
    native = NativeSpeechEncoder(enc, _Launcher(), th.device("cuda", 0))
    g = th.Generator(device="cuda").manual_seed(n + wav_len)
    noise = th.randn(n, wav_len, device="cuda", generator=g)
    t = th.arange(wav_len, device="cuda") / 16000.0
    voiced = sum(10.0 ** (-k / 2.0) * th.sin(2 * math.pi * 110.0 * (k + 1) * t * (1 + 0.01 * k)) for k in range(9))
    voiced = voiced[None] * (1 + 0.5 * th.sin(2 * math.pi * 3.0 * t))[None] + 1e-4 * noise
    for wav in (noise, voiced.expand(n, -1).contiguous()):
        prev = th.backends.cuda.matmul.allow_tf32
        th.backends.cuda.matmul.allow_tf32 = False
        try:
            with th.no_grad():
                raw = enc.wav2spec(wav) + 1e-6
                want = enc.mel_spec_norm(raw)
        finally:
            th.backends.cuda.matmul.allow_tf32 = prev
        got = native._mel(wav)
        assert got.shape == want.shape == (n, 128, wav_len // 512 + 1)
        assert ((got - want).norm() / want.norm()).item() < 1e-4
        assert (got - want).abs().max().item() < 5e-3   # unit-variance rows
        # the un-normalised mel powers, checked through the reference statement used on CPU
        raw2 = ref.mel_power_ref(wav, native.window, native.fb, native.preemph, 1e-6)
        assert ((raw2 - raw).norm() / raw.norm()).item() < 1e-5


@pytest.mark.parametrize("name", ["beat", "tedexp"])
def test_native_encoder_against_reference_golden_features(gd, name):
    """The CUDA encoder against the features the unmodified reference produced (tests/golden/*.npz)."""
    from gesture_b200.engine import _Launcher
    from gesture_b200.speech_native import NativeSpeechEncoder
    from util import build, load_golden, rel_l2, synthetic_wav
    g = load_golden(name)
    for weights in ("init", "boost"):
        model, _, _, _, L, _ = build(name, weights, device="cuda")
        native = NativeSpeechEncoder(model.speech_encoder, _Launcher(), th.device("cuda", 0))
        out = native(synthetic_wav(2, L, seed=123).cuda())
        for nm, f in zip(("low", "mid", "high"), out):
            err = rel_l2(f, g[f"{weights}.feat_{nm}"])
            assert err < 4e-3, (name, weights, nm, err)


def test_workspace_survives_changing_batch_sizes(gd):
    """One workspace (sized for the largest chunk) serves chunks of any smaller size in any order: a 5-clip chunk, a 2-clip
    tail, then the same again must reproduce the first pass bit for bit (the SE gate keeps arrival counters in scratch)."""
    from gesture_b200.engine import _Launcher
    from gesture_b200.modules import SpeechEncoder
    from gesture_b200.speech_native import NativeSpeechEncoder
    th.manual_seed(0)
    enc = SpeechEncoder(256).eval().cuda()
    native = NativeSpeechEncoder(enc, _Launcher(), th.device("cuda", 0), chunk=5)
    wav = th.randn(7, 16000, device="cuda", generator=th.Generator(device="cuda").manual_seed(9))
    first = native(wav)
    second = native(wav)
    third = native(wav[:5])
    for a, b, c in zip(first, second, third):
        assert th.isfinite(a).all()
        assert th.equal(a, b) and th.equal(a[:5], c)
