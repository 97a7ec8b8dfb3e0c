"""Host logic of the native speech encoder (gesture_b200/speech_native.py) without a GPU: the torch statements of the
speech kernels (tests/speech_ref.py) stand in for libgd_b200.so, and the three feature sequences must match the fp32
`SpeechEncoder` module (the restated HA2GSpeechEncoder, golden-checked against the reference in test_oracle_golden.py)
up to the bf16 storage of the feature maps."""
import pytest
import torch as th

import gesture_b200  # noqa: F401
from gesture_b200 import speech_native
from gesture_b200.modules import SpeechEncoder

from speech_ref import FakeLauncher


def randomise_batchnorm(enc, seed=1):
    g = th.Generator().manual_seed(seed)
    for m in enc.modules():
        if isinstance(m, th.nn.BatchNorm2d):
            m.running_mean.normal_(0, 0.1, generator=g)
            m.running_var.uniform_(0.5, 1.5, generator=g)
            m.weight.data.uniform_(0.8, 1.2, generator=g)
            m.bias.data.normal_(0, 0.1, generator=g)


# feature-map storage error of the two precisions on a default-init encoder (the boosted parity weights amplify it ~10x)
TOL = {"bf16x3": 2e-3, "bf16": 2e-2}


@pytest.mark.parametrize("precision", ["bf16x3", "bf16"])
@pytest.mark.parametrize("n,frames,chunk", [(3, 15, 2), (1, 10, 64)])
def test_native_plan_matches_module_on_cpu(monkeypatch, n, frames, chunk, precision):
    th.manual_seed(0)
    enc = SpeechEncoder(64).eval()
    randomise_batchnorm(enc)
    fake = FakeLauncher()
    monkeypatch.setattr(speech_native, "_p", fake.track)
    native = speech_native.NativeSpeechEncoder(enc, fake, th.device("cpu"), chunk=chunk, precision=precision)
    wav = th.randn(n, 512 * (frames - 1), generator=th.Generator().manual_seed(5))
    with th.no_grad():
        ref = enc(wavform=wav)
    out = native(wav)
    widths = [frames]
    for _ in range(3):
        widths.append((widths[-1] - 1) // 2 + 1)
    assert [o.shape[1] for o in out] == [widths[1] - 1, 2 * widths[2] - 2, 4 * widths[3] - 2]
    for a, b in zip(ref, out):
        assert a.shape == b.shape
        rel = ((a - b).norm() / a.norm()).item()
        assert rel < TOL[precision], (precision, rel)
    # 16 blocks x (2 convs + gate + tail) + 3 down-sample convs + 3 head convs + 2 shuffles + stem + 3 merged Linear layers
    assert fake.calls.count("gd_mel_power") == 1 and fake.calls.count("gd_instance_norm_rows") == 1
    per_chunk = {"gd_conv_taps_bf16": 38, "gd_se_gate": 16, "gd_se_residual_relu": 16, "gd_pixel_shuffle_rows": 2,
                 "gd_speech_stem": 1, "gd_linear_bf16": 3}
    chunks = (n + chunk - 1) // chunk
    for name, cnt in per_chunk.items():
        assert fake.calls.count(name) == cnt * chunks, name


def test_native_features_do_not_depend_on_the_batch(monkeypatch):
    th.manual_seed(0)
    enc = SpeechEncoder(64).eval()
    fake = FakeLauncher()
    monkeypatch.setattr(speech_native, "_p", fake.track)
    native = speech_native.NativeSpeechEncoder(enc, fake, th.device("cpu"), chunk=2)
    wav = th.randn(3, 512 * 9, generator=th.Generator().manual_seed(6))
    full = native(wav)
    last = native(wav[2:3])
    for a, b in zip(full, last):
        assert th.allclose(a[2:3], b, rtol=0, atol=1e-5)  # CPU matmul blocking differs with the row count; same plan


def test_head_merge_is_fc_then_projection():
    th.manual_seed(0)
    enc = SpeechEncoder(64).eval()
    r = enc.wav_encoder.feat_extractor
    hd = speech_native._Head(r.conv_mid, r.bn_mid, r.fc_mid, enc.wav_proj_layer, 2, 64, 62, 0)
    feat = th.randn(5, 32, 62)                      # (frames, c, y) as the reference flattens it (c*62 + y)
    ref = enc.wav_proj_layer(r.fc_mid(feat.reshape(5, -1)))
    mine = th.zeros(5, 62, 64)
    mine[:, :, :32] = feat.permute(0, 2, 1)          # our [y][c] layout, channels padded to 64
    out = mine.reshape(5, -1) @ hd.w.float().T + hd.b
    assert ((ref - out).norm() / ref.norm()).item() < 5e-3  # bf16 merged weight


@pytest.mark.parametrize("ci,co,k", [(32, 32, 3), (32, 64, 1), (64, 64, 3), (128, 64, 3)])
def test_split_weight_packings_agree(ci, co, k):
    """The three K layouts of a split-precision convolution (generic wrap walk, walk 1 / walk 2 with operand reuse) are the
    same arithmetic: hi*Whi + lo*Whi + hi*Wlo, which matches the float64 convolution to ~1e-5."""
    from speech_ref import conv_taps_ref, join
    g = th.Generator().manual_seed(ci + co + k)
    n, H, W = 2, 6, 5
    x = th.zeros(n, H + 2, W + 2, ci)
    x[:, 1:-1, 1:-1] = th.randn(n, H, W, ci, generator=g)
    x = x.reshape(-1, ci)
    hi = x.to(th.bfloat16)
    rows = th.cat([hi, (x - hi.float()).to(th.bfloat16)], dim=1).float()
    w4 = th.randn(co, ci, k, k, generator=g) / (ci * k * k) ** 0.5
    co_pad = speech_native._pad_to(co)
    gw = W + 2
    taps = [0] if k == 1 else [(ky - 1) * gw + (kx - 1) for ky in range(3) for kx in range(3)]
    scale, shift = th.ones(co_pad), th.zeros(co_pad)
    outs = []
    for reuse in (False, True):
        Wp, kpt, walk = speech_native._pack_conv(w4, ci, co_pad, 1, reuse)
        assert walk == (0 if not reuse else 1 if ci == 32 else 2) and Wp.shape == (co_pad, len(taps) * kpt)
        out = th.zeros(n * (H + 2) * gw, 2 * co)
        conv_taps_ref(rows, Wp, n, H + 2, gw, taps, kpt, None, scale, shift, 0, (1, H, 1, W), 1, out,
                      ((H + 2) * gw, gw, 1, gw + 1), co, 1, walk)
        outs.append(join(out, co, 1).view(n, H + 2, gw, co)[:, 1:-1, 1:-1])
    want = th.nn.functional.conv2d(join(rows, ci, 1).view(n, H + 2, gw, ci)[:, 1:-1, 1:-1].permute(0, 3, 1, 2).double(),
                                   w4.double(), padding=k // 2).permute(0, 2, 3, 1)
    for o in outs:
        assert ((o.double() - want).norm() / want.norm()).item() < 3e-5
    assert ((outs[0] - outs[1]).norm() / outs[0].norm()).item() < 1e-5  # same products, different fp32 summation order


@pytest.mark.parametrize("name", ["beat", "tedexp"])
def test_native_plan_against_reference_golden_features(monkeypatch, name):
    """Pinned to the real reference: the golden speech features were written by the unmodified HA2GSpeechEncoder
    (tests/golden/make_golden.py) on the same weights and synthetic wav."""
    from util import build, load_golden, rel_l2, synthetic_wav
    g = load_golden(name)
    for weights in ("init", "boost"):
        model, _, _, _, L, _ = build(name, weights)
        fake = FakeLauncher()
        monkeypatch.setattr(speech_native, "_p", fake.track)
        native = speech_native.NativeSpeechEncoder(model.speech_encoder, fake, th.device("cpu"))
        out = native(synthetic_wav(2, L, seed=123))
        for nm, f in zip(("low", "mid", "high"), out):
            assert rel_l2(f, g[f"{weights}.feat_{nm}"]) < 4e-3, (name, weights, nm, rel_l2(f, g[f"{weights}.feat_{nm}"]))
