"""TEST INFRASTRUCTURE (oracle) -- not product code.  PARITY UNPINNED at the demucs boundary.

CPU restatement of the parts of the third-party package ``demucs==4.0.1``
(reference `requirements.txt:1`) that the hot path touches.  The package source
is NOT vendored under /root/reference and is not installable here (no network),
and the reference repository holds no golden vectors for it, so this file is a
restatement of the *published* demucs 4.0.1 algorithm (htdemucs.py, hdemucs.py,
demucs.py, transformer.py, spec.py) for the pretrained ``htdemucs`` signature
``955717e8`` configuration.  Structure (module names, parameter shapes) is
checked against the reference's own dumps:
  /root/reference/src/models/stem_separation/AudioTextHTDemucs_Full.txt:3-629
  /root/reference/src/models/stem_separation/HTDemucs_Fwd_Pass.txt:4-89,147
Call sites in the reference that define which members must exist:
  /root/reference/src/models/stem_separation/ATHTDemucs_v2.py:171,197-234,261-262,310

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this module.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

NFFT = 4096
HOP = 1024


# --------------------------------------------------------------------------- spec.py
def spectro(x: torch.Tensor, n_fft: int = NFFT, hop_length: int = HOP) -> torch.Tensor:
    """demucs/spec.py:spectro -- normalized, centred, reflect-padded hann STFT."""
    *other, length = x.shape
    x = x.reshape(-1, length)
    z = torch.stft(
        x, n_fft, hop_length, window=torch.hann_window(n_fft).to(x), win_length=n_fft,
        normalized=True, center=True, return_complex=True, pad_mode="reflect",
    )
    _, freqs, frames = z.shape
    return z.view(*other, freqs, frames)


def ispectro(z: torch.Tensor, hop_length: int = HOP, length: Optional[int] = None) -> torch.Tensor:
    """demucs/spec.py:ispectro."""
    *other, freqs, frames = z.shape
    n_fft = 2 * freqs - 2
    z = z.view(-1, freqs, frames)
    x = torch.istft(
        z, n_fft, hop_length, window=torch.hann_window(n_fft).to(z.real), win_length=n_fft,
        normalized=True, length=length, center=True,
    )
    _, length = x.shape
    return x.view(*other, length)


def pad1d(x: torch.Tensor, paddings, mode: str = "constant", value: float = 0.0) -> torch.Tensor:
    """demucs/hdemucs.py:pad1d -- reflect pad that zero-extends inputs shorter than the pad."""
    length = x.shape[-1]
    padding_left, padding_right = paddings
    if mode == "reflect":
        max_pad = max(padding_left, padding_right)
        if length <= max_pad:
            extra_pad = max_pad - length + 1
            extra_pad_right = min(padding_right, extra_pad)
            extra_pad_left = extra_pad - extra_pad_right
            paddings = (padding_left - extra_pad_left, padding_right - extra_pad_right)
            x = F.pad(x, (extra_pad_left, extra_pad_right))
    out = F.pad(x, paddings, mode, value)
    assert out.shape[-1] == length + padding_left + padding_right
    return out


# --------------------------------------------------------------------------- demucs.py
class LayerScale(nn.Module):
    """demucs LayerScale: per-channel learnt scale, channel-first unless channel_last."""

    def __init__(self, channels: int, init: float = 0.0, channel_last: bool = False):
        super().__init__()
        self.channel_last = channel_last
        self.scale = nn.Parameter(torch.zeros(channels))
        self.scale.data[:] = init

    def forward(self, x):
        if self.channel_last:
            return self.scale * x
        return self.scale[:, None] * x


class DConv(nn.Module):
    """demucs/demucs.py:DConv (depth 2, compress 8, dilated k3, GroupNorm(1), GELU, GLU, LayerScale)."""

    def __init__(self, channels: int, compress: float = 8, depth: int = 2, init: float = 1e-3):
        super().__init__()
        hidden = int(channels / compress)
        self.layers = nn.ModuleList()
        for d in range(depth):
            dilation = 2 ** d
            self.layers.append(nn.Sequential(
                nn.Conv1d(channels, hidden, 3, dilation=dilation, padding=dilation),
                nn.GroupNorm(1, hidden),
                nn.GELU(),
                nn.Conv1d(hidden, 2 * channels, 1),
                nn.GroupNorm(1, 2 * channels),
                nn.GLU(1),
                LayerScale(channels, init),
            ))

    def forward(self, x):
        for layer in self.layers:
            x = x + layer(x)
        return x


# --------------------------------------------------------------------------- hdemucs.py
class ScaledEmbedding(nn.Module):
    """demucs/hdemucs.py:ScaledEmbedding(512, 48, scale=10, smooth=True)."""

    def __init__(self, num_embeddings: int, embedding_dim: int, scale: float = 10.0, smooth: bool = False):
        super().__init__()
        self.embedding = nn.Embedding(num_embeddings, embedding_dim)
        if smooth:
            weight = torch.cumsum(self.embedding.weight.data, dim=0)
            weight = weight / torch.arange(1, num_embeddings + 1).to(weight).sqrt()[:, None]
            self.embedding.weight.data[:] = weight
        self.embedding.weight.data /= scale
        self.scale = scale

    def forward(self, x):
        return self.embedding(x) * self.scale


class HEncLayer(nn.Module):
    """demucs/hdemucs.py:HEncLayer with norm=False (Identity norms), dconv, rewrite, pad."""

    def __init__(self, chin: int, chout: int, kernel_size: int = 8, stride: int = 4, freq: bool = True):
        super().__init__()
        self.freq = freq
        self.kernel_size = kernel_size
        self.stride = stride
        self.empty = False
        pad = kernel_size // 4
        if freq:
            self.conv = nn.Conv2d(chin, chout, (kernel_size, 1), (stride, 1), (pad, 0))
        else:
            self.conv = nn.Conv1d(chin, chout, kernel_size, stride, pad)
        self.norm1 = nn.Identity()
        if freq:
            self.rewrite = nn.Conv2d(chout, 2 * chout, 1, 1, 0)
        else:
            self.rewrite = nn.Conv1d(chout, 2 * chout, 1, 1, 0)
        self.norm2 = nn.Identity()
        self.dconv = DConv(chout)

    def forward(self, x, inject=None):
        if not self.freq and x.dim() == 4:
            B, C, Fr, T = x.shape
            x = x.view(B, -1, T)
        if not self.freq:
            le = x.shape[-1]
            if not le % self.stride == 0:
                x = F.pad(x, (0, self.stride - (le % self.stride)))
        y = self.conv(x)
        if inject is not None:
            assert inject.shape[-1] == y.shape[-1], (inject.shape, y.shape)
            if inject.dim() == 3 and y.dim() == 4:
                inject = inject[:, :, None]
            y = y + inject
        y = F.gelu(self.norm1(y))
        if self.freq:
            B, C, Fr, T = y.shape
            y = y.permute(0, 2, 1, 3).reshape(-1, C, T)
        y = self.dconv(y)
        if self.freq:
            y = y.view(B, Fr, C, T).permute(0, 2, 1, 3)
        z = self.norm2(self.rewrite(y))
        return F.glu(z, dim=1)


class _DeadDecLayer(nn.Module):
    """Parameter container for the HTDemucs 4-stem decoder layers, which the hot
    path never calls (reference AudioTextHTDemucs_Full.txt:118-231, 346-459); kept so
    that checkpoints carrying these keys exercise load_state_dict(strict=False)."""

    def __init__(self, chin: int, chout: int, freq: bool):
        super().__init__()
        if freq:
            self.conv_tr = nn.ConvTranspose2d(chin, chout, (8, 1), (4, 1))
            self.rewrite = nn.Conv2d(chin, 2 * chin, 3, 1, 1)
        else:
            self.conv_tr = nn.ConvTranspose1d(chin, chout, 8, 4)
            self.rewrite = nn.Conv1d(chin, 2 * chin, 3, 1, 1)
        self.dconv = DConv(chin)


# --------------------------------------------------------------------------- transformer.py
def create_sin_embedding(length: int, dim: int, shift: int = 0, max_period: float = 10000.0):
    pos = shift + torch.arange(length).view(-1, 1, 1)
    half_dim = dim // 2
    adim = torch.arange(dim // 2).view(1, 1, -1)
    phase = pos / (max_period ** (adim / (half_dim - 1)))
    return torch.cat([torch.cos(phase), torch.sin(phase)], dim=-1)  # (T, 1, C)


def create_2d_sin_embedding(d_model: int, height: int, width: int, max_period: float = 10000.0):
    pe = torch.zeros(d_model, height, width)
    d = d_model // 2
    div_term = torch.exp(torch.arange(0.0, d, 2) * -(math.log(max_period) / d))
    pos_w = torch.arange(0.0, width).unsqueeze(1)
    pos_h = torch.arange(0.0, height).unsqueeze(1)
    pe[0:d:2, :, :] = torch.sin(pos_w * div_term).transpose(0, 1).unsqueeze(1).repeat(1, height, 1)
    pe[1:d:2, :, :] = torch.cos(pos_w * div_term).transpose(0, 1).unsqueeze(1).repeat(1, height, 1)
    pe[d::2, :, :] = torch.sin(pos_h * div_term).transpose(0, 1).unsqueeze(2).repeat(1, 1, width)
    pe[d + 1::2, :, :] = torch.cos(pos_h * div_term).transpose(0, 1).unsqueeze(2).repeat(1, 1, width)
    return pe[None, :]  # (1, C, Fr, T1)


class MyGroupNorm(nn.GroupNorm):
    """GroupNorm over (tokens x channels) of a (B, T, C) tensor."""

    def forward(self, x):
        x = x.transpose(1, 2)
        return super().forward(x).transpose(1, 2)


class MyTransformerEncoderLayer(nn.Module):
    """Pre-norm self-attention layer with LayerScale and MyGroupNorm norm_out."""

    def __init__(self, d_model: int = 512, nhead: int = 8, dim_feedforward: int = 2048,
                 dropout: float = 0.02, init_values: float = 1e-4):
        super().__init__()
        self.self_attn = nn.MultiheadAttention(d_model, nhead, dropout=dropout, batch_first=True)
        self.linear1 = nn.Linear(d_model, dim_feedforward)
        self.dropout = nn.Dropout(dropout)
        self.linear2 = nn.Linear(dim_feedforward, d_model)
        self.norm1 = nn.LayerNorm(d_model, eps=1e-5)
        self.norm2 = nn.LayerNorm(d_model, eps=1e-5)
        self.dropout1 = nn.Dropout(dropout)
        self.dropout2 = nn.Dropout(dropout)
        self.norm_out = MyGroupNorm(1, d_model)
        self.gamma_1 = LayerScale(d_model, init_values, True)
        self.gamma_2 = LayerScale(d_model, init_values, True)

    def forward(self, x):
        h = self.norm1(x)
        h = self.self_attn(h, h, h, need_weights=False)[0]
        x = x + self.gamma_1(self.dropout1(h))
        h = self.linear2(self.dropout(F.gelu(self.linear1(self.norm2(x)))))
        x = x + self.gamma_2(self.dropout2(h))
        return self.norm_out(x)


class CrossTransformerEncoderLayer(nn.Module):
    """Pre-norm cross-attention layer (q from own branch, k=v from the other)."""

    def __init__(self, d_model: int = 512, nhead: int = 8, dim_feedforward: int = 2048,
                 dropout: float = 0.02, init_values: float = 1e-4):
        super().__init__()
        self.cross_attn = nn.MultiheadAttention(d_model, nhead, dropout=dropout, batch_first=True)
        self.linear1 = nn.Linear(d_model, dim_feedforward)
        self.dropout = nn.Dropout(dropout)
        self.linear2 = nn.Linear(dim_feedforward, d_model)
        self.norm1 = nn.LayerNorm(d_model, eps=1e-5)
        self.norm2 = nn.LayerNorm(d_model, eps=1e-5)
        self.norm3 = nn.LayerNorm(d_model, eps=1e-5)
        self.norm_out = MyGroupNorm(1, d_model)
        self.gamma_1 = LayerScale(d_model, init_values, True)
        self.gamma_2 = LayerScale(d_model, init_values, True)
        self.dropout1 = nn.Dropout(dropout)
        self.dropout2 = nn.Dropout(dropout)

    def forward(self, q, k):
        kn = self.norm2(k)
        h = self.cross_attn(self.norm1(q), kn, kn, need_weights=False)[0]
        x = q + self.gamma_1(self.dropout1(h))
        h = self.linear2(self.dropout(F.gelu(self.linear1(self.norm3(x)))))
        x = x + self.gamma_2(self.dropout2(h))
        return self.norm_out(x)


class CrossTransformerEncoder(nn.Module):
    """demucs/transformer.py:CrossTransformerEncoder, 5 layers, sin embeddings, cross_first=False."""

    def __init__(self, dim: int = 512, num_heads: int = 8, num_layers: int = 5, hidden_scale: float = 4.0,
                 dropout: float = 0.02, max_period: float = 10000.0, weight_pos_embed: float = 1.0):
        super().__init__()
        hidden = int(dim * hidden_scale)
        self.num_layers = num_layers
        self.classic_parity = 0
        self.max_period = max_period
        self.weight_pos_embed = weight_pos_embed
        self.norm_in = nn.LayerNorm(dim)
        self.norm_in_t = nn.LayerNorm(dim)
        self.layers = nn.ModuleList()
        self.layers_t = nn.ModuleList()
        for idx in range(num_layers):
            if idx % 2 == self.classic_parity:
                self.layers.append(MyTransformerEncoderLayer(dim, num_heads, hidden, dropout))
                self.layers_t.append(MyTransformerEncoderLayer(dim, num_heads, hidden, dropout))
            else:
                self.layers.append(CrossTransformerEncoderLayer(dim, num_heads, hidden, dropout))
                self.layers_t.append(CrossTransformerEncoderLayer(dim, num_heads, hidden, dropout))

    def forward(self, x, xt):
        B, C, Fr, T1 = x.shape
        pe2 = create_2d_sin_embedding(C, Fr, T1, self.max_period).to(x)
        pe2 = pe2.permute(0, 3, 2, 1).reshape(1, T1 * Fr, C)          # b (t1 fr) c
        x = x.permute(0, 3, 2, 1).reshape(B, T1 * Fr, C)              # b (t1 fr) c
        x = self.norm_in(x)
        x = x + self.weight_pos_embed * pe2
        B, C, T2 = xt.shape
        xt = xt.permute(0, 2, 1)
        pe1 = create_sin_embedding(T2, C, 0, self.max_period).to(xt).permute(1, 0, 2)  # 1 t2 c
        xt = self.norm_in_t(xt)
        xt = xt + self.weight_pos_embed * pe1
        for idx in range(self.num_layers):
            if idx % 2 == self.classic_parity:
                x = self.layers[idx](x)
                xt = self.layers_t[idx](xt)
            else:
                old_x = x
                x = self.layers[idx](x, xt)
                xt = self.layers_t[idx](xt, old_x)
        x = x.reshape(B, T1, Fr, C).permute(0, 3, 2, 1)
        xt = xt.permute(0, 2, 1)
        return x, xt


# --------------------------------------------------------------------------- htdemucs.py
def rescale_module(module: nn.Module, reference: float = 0.1) -> None:
    """demucs/demucs.py:rescale_module -- called at the end of HTDemucs.__init__."""
    for sub in module.modules():
        if isinstance(sub, (nn.Conv1d, nn.ConvTranspose1d, nn.Conv2d, nn.ConvTranspose2d)):
            std = sub.weight.std().detach()
            scale = (std / reference) ** 0.5
            sub.weight.data /= scale
            if sub.bias is not None:
                sub.bias.data /= scale


class HTDemucs(nn.Module):
    """The members of demucs.htdemucs.HTDemucs that ATHTDemucs_v2.py reads, for the
    pretrained configuration (sources=4, audio_channels=2, channels=48, growth=2,
    nfft=4096, cac=True, depth=4, rewrite, freq_emb=0.2, emb_smooth, kernel 8 stride 4,
    dconv_comp=8, dconv_init=1e-3, norm_starts=4, bottom_channels=512, t_layers=5)."""

    def __init__(self, include_dead_decoders: bool = True):
        super().__init__()
        self.audio_channels = 2
        self.nfft = NFFT
        self.hop_length = HOP
        self.cac = True
        self.depth = 4
        self.bottom_channels = 512
        self.freq_emb_scale = 0.2
        chans = [4, 48, 96, 192, 384]
        self.encoder = nn.ModuleList()
        self.decoder = nn.ModuleList()
        self.tencoder = nn.ModuleList()
        self.tdecoder = nn.ModuleList()
        for i in range(4):
            self.encoder.append(HEncLayer(chans[i], chans[i + 1], freq=True))
            self.tencoder.append(HEncLayer(2 if i == 0 else chans[i], chans[i + 1], freq=False))
        if include_dead_decoders:
            outs = [16, 48, 96, 192]  # sources * (cac channels) at the last layer
            for i in reversed(range(4)):
                self.decoder.append(_DeadDecLayer(chans[i + 1], outs[i], freq=True))
                self.tdecoder.append(_DeadDecLayer(chans[i + 1], 8 if i == 0 else chans[i], freq=False))
        self.freq_emb = ScaledEmbedding(512, 48, smooth=True, scale=10.0)
        self.channel_upsampler = nn.Conv1d(384, 512, 1)
        self.channel_downsampler = nn.Conv1d(512, 384, 1)
        self.channel_upsampler_t = nn.Conv1d(384, 512, 1)
        self.channel_downsampler_t = nn.Conv1d(512, 384, 1)
        self.crosstransformer = CrossTransformerEncoder()
        rescale_module(self, 0.1)

    def _spec(self, x):
        hl = self.hop_length
        le = int(math.ceil(x.shape[-1] / hl))
        pad = hl // 2 * 3
        x = pad1d(x, (pad, pad + le * hl - x.shape[-1]), mode="reflect")
        z = spectro(x, self.nfft, hl)[..., :-1, :]
        assert z.shape[-1] == le + 4, (z.shape, x.shape, le)
        return z[..., 2:2 + le]

    def _ispec(self, z, length=None, scale=0):
        hl = self.hop_length // (4 ** scale)
        z = F.pad(z, (0, 0, 0, 1))
        z = F.pad(z, (2, 2))
        pad = hl // 2 * 3
        le = hl * int(math.ceil(length / hl)) + 2 * pad
        x = ispectro(z, hl, length=le)
        return x[..., pad:pad + length]

    def _magnitude(self, z):
        B, C, Fr, T = z.shape
        m = torch.view_as_real(z).permute(0, 1, 4, 2, 3)
        return m.reshape(B, C * 2, Fr, T)
