"""Parameter containers for the denoiser (same `state_dict` key set and shapes as the reference, so a reference
checkpoint loads with `strict=True`) and the loop-invariant speech-conditioning producer.

Only the speech encoder has a torch `forward` here: it runs ONCE per clip (the reference re-runs it at every one
of the 1000 steps, models/model.py:54-56,94-96) and stays on torch/cuDNN.  The transformer decoder has no torch
forward at all — its math lives in the sm_100a kernels driven by `engine.py`; these modules only own the
parameters (`models/nn.py`, `models/modules/transformer.py` define the layouts they mirror).

Construction order follows the reference constructors so that the same `torch.manual_seed` yields the same
random initialisation (checked in tests/golden/make_golden.py against the real reference).
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F


# ------------------------------------------------------------------------------------------ decoder parameters
class _HeadProjection(nn.Module):
    """`PrepareForMultiHeadAttention` parameters (transformer.py:47-59): key `...linear.{weight,bias}`."""

    def __init__(self, d_model, heads, d_k):
        super().__init__()
        self.linear = nn.Linear(d_model, heads * d_k, bias=True)


class _TokenConv(nn.Module):
    """`SpatialDepthWiseConv` parameters (transformer.py:19-44): key `...conv.{weight (d_k,1,3), bias}`."""

    def __init__(self, d_k):
        super().__init__()
        self.conv = nn.Conv1d(d_k, d_k, kernel_size=(3,), padding=(2,), groups=d_k)


class DConvAttentionParams(nn.Module):
    """`MultiDConvHeadAttention` (transformer.py:62-126): query/key/value = Sequential(projection, token conv)."""

    def __init__(self, heads, d_model):
        super().__init__()
        assert d_model % heads == 0, "d_model msut be divisible by heads."
        self.heads, self.d_k = heads, d_model // heads
        # registration order query, key, value, output (it fixes parameters() order, hence the xavier RNG order)
        self.query = _HeadProjection(d_model, heads, self.d_k)
        self.key = _HeadProjection(d_model, heads, self.d_k)
        self.value = _HeadProjection(d_model, heads, self.d_k)
        self.output = nn.Linear(d_model, d_model)
        self.query = nn.Sequential(self.query, _TokenConv(self.d_k))
        self.key = nn.Sequential(self.key, _TokenConv(self.d_k))
        self.value = nn.Sequential(self.value, _TokenConv(self.d_k))


class FeedForwardParams(nn.Module):
    """`FeedForward` (transformer.py:129-154), d_ff = 4 d_model, squared-ReLU in between."""

    def __init__(self, d_model):
        super().__init__()
        self.layer1 = nn.Linear(d_model, 4 * d_model)
        self.layer2 = nn.Linear(4 * d_model, d_model)


def _xavier_matrices(module):
    for p in module.parameters():
        if p.dim() > 1:
            nn.init.xavier_uniform_(p)


class JointLayerParams(nn.Module):
    """`CrossAttentionLayer` (nn.py:55-125): pose stream, memory stream, joint attention, two FFNs."""

    def __init__(self, d_model, heads, with_memory_ffn):
        super().__init__()
        sa, sam, ca = (DConvAttentionParams(heads, d_model) for _ in range(3))
        ff = FeedForwardParams(d_model)
        ffm = FeedForwardParams(d_model) if with_memory_ffn else None
        self.size = d_model
        self.norm_self_attn = nn.LayerNorm([d_model])
        self.self_attn = sa
        self.norm_self_attn_mem = nn.LayerNorm([d_model])
        self.self_attn_mem = sam
        self.norm_cross_attn = nn.LayerNorm([d_model])
        self.cross_attn = ca
        self.norm_ff = nn.LayerNorm([d_model])
        self.feed_forward = ff
        self.feed_forward_mem = ffm
        if ffm is not None:
            self.norm_ff_mem = nn.LayerNorm([d_model])
        _xavier_matrices(self)


class OnewayLayerParams(nn.Module):
    """`OnewayCrossAttentionLayer` (nn.py:128-174)."""

    def __init__(self, d_model, heads):
        super().__init__()
        sa, ca = DConvAttentionParams(heads, d_model), DConvAttentionParams(heads, d_model)
        ff = FeedForwardParams(d_model)
        self.size = d_model
        self.norm_self_attn = nn.LayerNorm([d_model])
        self.self_attn = sa
        self.norm_cross_attn = nn.LayerNorm([d_model])
        self.cross_attn = ca
        self.norm_ff = nn.LayerNorm([d_model])
        self.feed_forward = ff
        _xavier_matrices(self)


class PoseDecoderParams(nn.Module):
    """`CrossAttention` (nn.py:381-447, kind='cross_attention') / `OnewayCrossAttention` (nn.py:177-228)."""

    def __init__(self, kind, d_x, d_model, heads, n_layers, d_out):
        super().__init__()
        self.kind, self.heads, self.n_layers, self.d_model = kind, heads, n_layers, d_model
        self.emb_x = nn.Linear(d_x, d_model)
        self.emb_mem = nn.Linear(d_model, d_model)
        if kind == "cross_attention":
            layers = [JointLayerParams(d_model, heads, True) for _ in range(n_layers - 1)]
            layers.append(JointLayerParams(d_model, heads, False))  # last layer: no memory FFN (nn.py:411-418)
        elif kind == "oneway_cross_attention":
            layers = [OnewayLayerParams(d_model, heads) for _ in range(n_layers)]
        else:
            raise ValueError(f"Unsupported decoder type {kind}.")
        self.layers = nn.ModuleList(layers)
        self.out_layers = nn.Sequential(nn.LayerNorm([d_model]), nn.Linear(d_model, d_out))


class StepEncoderParams(nn.Module):
    """`DiffusionStepEncoder` (nn.py:38-52): Linear -> SiLU -> Linear (-> Dropout p=0)."""

    def __init__(self, d_model, dropout_rate):
        super().__init__()
        self.proj = nn.Sequential(nn.Linear(d_model, d_model), nn.SiLU(), nn.Linear(d_model, d_model),
                                  nn.Dropout(p=dropout_rate))
        self.d_model = d_model


def positional_table(d_model, length):
    """`get_positional_encoding` (transformer.py:157-166): sin on even, cos on odd features, fp32."""
    pe = torch.zeros(length, d_model)
    pos = torch.arange(0, length, dtype=torch.float32).unsqueeze(1)
    div = torch.exp(torch.arange(0, d_model, 2, dtype=torch.float32) * -(math.log(10000.0) / d_model))
    pe[:, 0::2] = torch.sin(pos * div)
    pe[:, 1::2] = torch.cos(pos * div)
    return pe


def step_embedding_table(d_model, n_steps, max_period=10000):
    """`diffusion_step_embedding` (nn.py:17-35) for t = 0..n_steps-1: [cos | sin]."""
    half = d_model // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32) / half)
    args = torch.arange(n_steps)[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


# ------------------------------------------------------------------------------------------ speech encoder
class _PreEmphasis(nn.Module):
    """y[t] = x[t] - 0.97 x[t-1] with reflect padding (ha2g/model/utils.py:22-37); buffer key `flipped_filter`."""

    def __init__(self, coef=0.97):
        super().__init__()
        self.register_buffer("flipped_filter", torch.FloatTensor([-coef, 1.0]).unsqueeze(0).unsqueeze(0))

    def forward(self, wav):
        assert wav.dim() == 2, "The number of dimensions of input tensor must be 2!"
        return F.conv1d(F.pad(wav.unsqueeze(1), (1, 0), "reflect"), self.flipped_filter).squeeze(1)


class _Stft(nn.Module):
    def __init__(self, n_fft, hop):
        super().__init__()
        self.n_fft, self.hop = n_fft, hop
        self.register_buffer("window", torch.hann_window(n_fft))

    def forward(self, x):
        s = torch.stft(x, self.n_fft, hop_length=self.hop, win_length=self.n_fft, window=self.window, center=True,
                       pad_mode="reflect", normalized=False, onesided=True, return_complex=True)
        return s.abs().pow(2.0)


class _MelScale(nn.Module):
    def __init__(self, n_mels, sample_rate, n_freqs):
        super().__init__()
        hz2mel = lambda f: 2595.0 * math.log10(1.0 + f / 700.0)  # noqa: E731  (HTK)
        freqs = torch.linspace(0, sample_rate // 2, n_freqs)
        m_pts = torch.linspace(hz2mel(0.0), hz2mel(sample_rate / 2.0), n_mels + 2)
        f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)
        f_diff = f_pts[1:] - f_pts[:-1]
        slopes = f_pts.unsqueeze(0) - freqs.unsqueeze(1)
        fb = torch.max(torch.zeros(1), torch.min((-1.0 * slopes[:, :-2]) / f_diff[:-1], slopes[:, 2:] / f_diff[1:]))
        self.register_buffer("fb", fb)

    def forward(self, spec):
        return torch.matmul(spec.transpose(-1, -2), self.fb).transpose(-1, -2)


class MelSpectrogram(nn.Module):
    """torchaudio.transforms.MelSpectrogram(16 kHz, n_fft 1024, hop 512, 128 mels) restated with the same
    persistent buffers (`spectrogram.window`, `mel_scale.fb`) so checkpoints keep loading (speech_encoder.py:18-25)."""

    def __init__(self, sample_rate=16000, n_fft=1024, hop_length=512, n_mels=128):
        super().__init__()
        self.spectrogram = _Stft(n_fft, hop_length)
        self.mel_scale = _MelScale(n_mels, sample_rate, n_fft // 2 + 1)

    def forward(self, wav):
        return self.mel_scale(self.spectrogram(wav))


class _SqueezeExcite(nn.Module):
    def __init__(self, channels, reduction=8):
        super().__init__()
        self.fc = nn.Sequential(nn.Linear(channels, channels // reduction), nn.ReLU(inplace=True),
                                nn.Linear(channels // reduction, channels), nn.Sigmoid())

    def forward(self, x):
        return x * self.fc(x.mean(dim=(2, 3)))[:, :, None, None]


class _SEBlock(nn.Module):
    """`SEBasicBlock` (ResNetBlocks.py:7-37): conv-relu-BN (in that order), conv-BN, SE, residual, relu."""

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 3, stride=stride, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.se = _SqueezeExcite(planes)
        self.downsample = downsample

    def forward(self, x):
        out = self.bn1(F.relu(self.conv1(x)))
        out = self.se(self.bn2(self.conv2(out)))
        return F.relu(out + (x if self.downsample is None else self.downsample(x)))


class _ResNetSE34(nn.Module):
    """`ResNetSE` with SEBasicBlock [3,4,6,3], filters 32/64/128/256 and the three pyramid heads
    (ResNetSE34V2.py:13-220).  Input (N,1,128,frames) -> three (N, T_k, 32) feature sequences."""

    def __init__(self, n_out=32):
        super().__init__()
        f = [32, 64, 128, 256]
        self._inplanes = f[0]
        self.conv1 = nn.Conv2d(1, f[0], 3, stride=1, padding=1)
        self.bn1 = nn.BatchNorm2d(f[0])
        self.conv_low, self.bn_low, self.fc_low = nn.Conv2d(64, 64, 2), nn.BatchNorm2d(64), nn.Linear(63 * 64, n_out)
        self.conv_mid, self.bn_mid, self.fc_mid = nn.Conv2d(32, 32, 3), nn.BatchNorm2d(32), nn.Linear(62 * 32, n_out)
        self.conv_high, self.bn_high, self.fc_high = nn.Conv2d(16, 16, 3), nn.BatchNorm2d(16), nn.Linear(62 * 16, n_out)
        self.layer1 = self._stage(f[0], 3, 1)
        self.layer2 = self._stage(f[1], 4, 2)
        self.layer3 = self._stage(f[2], 6, 2)
        self.layer4 = self._stage(f[3], 3, 2)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)

    def _stage(self, planes, blocks, stride):
        down = None
        if stride != 1 or self._inplanes != planes:
            down = nn.Sequential(nn.Conv2d(self._inplanes, planes, 1, stride=stride, bias=False), nn.BatchNorm2d(planes))
        seq = [_SEBlock(self._inplanes, planes, stride, down)]
        self._inplanes = planes
        seq += [_SEBlock(planes, planes) for _ in range(1, blocks)]
        return nn.Sequential(*seq)

    @staticmethod
    def _head(feat, conv, bn, fc, shuffle):
        if shuffle > 1:
            feat = F.pixel_shuffle(feat, shuffle)
        feat = bn(F.relu(conv(feat)))
        n = feat.shape[0]
        return fc(feat.reshape(n, -1, feat.shape[-1]).transpose(1, 2))

    def forward(self, x):
        x = self.layer1(self.bn1(F.relu(self.conv1(x))))
        f1 = self.layer2(x)
        f2 = self.layer3(f1)
        f3 = self.layer4(f2)
        return (self._head(f1, self.conv_low, self.bn_low, self.fc_low, 1),
                self._head(f2, self.conv_mid, self.bn_mid, self.fc_mid, 2),
                self._head(f3, self.conv_high, self.bn_high, self.fc_high, 4))


class _WavEncoder(nn.Module):
    """`Hierarchical_WavEncoder` (hierarchy_net.py:10-19): key prefix `feat_extractor.`"""

    def __init__(self):
        super().__init__()
        self.feat_extractor = _ResNetSE34(32)

    def forward(self, mel):
        return self.feat_extractor(mel.unsqueeze(1))


class SpeechEncoder(nn.Module):
    """`HA2GSpeechEncoder` (speech_encoder.py:9-61): wav (N, T_wav) -> three (N, T_k, d_model) feature sequences."""

    def __init__(self, d_model, dropout_prob=0.0):
        super().__init__()
        self.wav2spec = nn.Sequential(_PreEmphasis(), MelSpectrogram())
        self.wav2spec.requires_grad_(False)
        self.mel_spec_norm = nn.InstanceNorm1d(128)
        self.wav_encoder = _WavEncoder()
        self.wav_proj_layer = nn.Linear(32, d_model)
        self.dropout = nn.Dropout(p=dropout_prob)

    def forward(self, *, wavform):
        mel = self.mel_spec_norm(self.wav2spec(wavform) + 1e-6)
        low, mid, high = self.wav_encoder(mel)
        return tuple(self.wav_proj_layer(self.dropout(f)) for f in (low, mid, high))
