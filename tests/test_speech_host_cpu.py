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
