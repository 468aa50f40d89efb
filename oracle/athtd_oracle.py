"""TEST INFRASTRUCTURE (oracle) -- not product code.  PARITY UNPINNED at the demucs boundary.

CPU fp32 functional restatement of the reference hot path

    AudioTextHTDemucs.forward      /root/reference/src/models/stem_separation/ATHTDemucs_v2.py:250-326
    AudioTextHTDemucs._encode      ATHTDemucs_v2.py:190-236
    TextCrossAttention             ATHTDemucs_v2.py:21-58
    FreqDecoder / TimeDecoder      ATHTDemucs_v2.py:61-104 / 107-139

written over a plain ``state_dict`` (same key layout as the reference module, SURVEY.md
Appendix E) instead of nn.Modules, with named taps after every stage.  The
``demucs==4.0.1`` pieces (HEncLayer, DConv, CrossTransformerEncoder, _spec/_ispec) follow
the published package algorithm (see oracle/demucs_shim.py header): the reference has no
golden vectors for them, hence "parity unpinned".  The AudioTextHTDemucs half is pinned
in this container by ``oracle/validate_against_reference.py``, which imports the real
reference file and checks this restatement against it; its outputs are committed as
fixtures under tests/golden/.

The CLAP text tower (ATHTDemucs_v2.py:238-248) is out of scope: the oracle takes the
(B, 512) text embedding directly.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this module.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

from .demucs_shim import (HOP, NFFT, create_2d_sin_embedding, create_sin_embedding, ispectro,
                          pad1d, spectro)

Taps = Optional[Dict[str, torch.Tensor]]


def _tap(taps: Taps, name: str, t: torch.Tensor) -> None:
    if taps is not None:
        taps[name] = t.detach().clone()


# ----------------------------------------------------------------------------- spectral
def spec(wav: torch.Tensor) -> torch.Tensor:
    """htdemucs._spec (call site ATHTDemucs_v2.py:261): [B,2,L] -> complex [B,2,2048,ceil(L/1024)]."""
    le = int(math.ceil(wav.shape[-1] / HOP))
    pad = HOP // 2 * 3
    x = pad1d(wav, (pad, pad + le * HOP - wav.shape[-1]), mode="reflect")
    z = spectro(x, NFFT, HOP)[..., :-1, :]
    assert z.shape[-1] == le + 4
    return z[..., 2:2 + le]


def magnitude(z: torch.Tensor) -> torch.Tensor:
    """htdemucs._magnitude with cac=True (call site ATHTDemucs_v2.py:262)."""
    B, C, Fr, T = z.shape
    return torch.view_as_real(z).permute(0, 1, 4, 2, 3).reshape(B, C * 2, Fr, T)


def ispec(z: torch.Tensor, length: int) -> torch.Tensor:
    """htdemucs._ispec (call site ATHTDemucs_v2.py:310)."""
    z = F.pad(z, (0, 0, 0, 1))
    z = F.pad(z, (2, 2))
    pad = HOP // 2 * 3
    le = HOP * int(math.ceil(length / HOP)) + 2 * pad
    x = ispectro(z, HOP, length=le)
    return x[..., pad:pad + length]


# ----------------------------------------------------------------------------- encoder
def dconv(sd, prefix: str, x: torch.Tensor) -> torch.Tensor:
    """demucs DConv, depth 2 (printed tree AudioTextHTDemucs_Full.txt:10-31). x: [N, C, T]."""
    for d in range(2):
        p = f"{prefix}.layers.{d}"
        dil = 2 ** d
        h = F.conv1d(x, sd[f"{p}.0.weight"], sd[f"{p}.0.bias"], dilation=dil, padding=dil)
        h = F.group_norm(h, 1, sd[f"{p}.1.weight"], sd[f"{p}.1.bias"], eps=1e-5)
        h = F.gelu(h)
        h = F.conv1d(h, sd[f"{p}.3.weight"], sd[f"{p}.3.bias"])
        h = F.group_norm(h, 1, sd[f"{p}.4.weight"], sd[f"{p}.4.bias"], eps=1e-5)
        h = F.glu(h, dim=1)
        x = x + sd[f"{p}.6.scale"][:, None] * h
    return x


def henc_layer(sd, prefix: str, x: torch.Tensor, freq: bool, taps: Taps = None) -> torch.Tensor:
    """demucs HEncLayer (norm=False -> Identity norms), SURVEY.md Appendix A2."""
    if freq:
        y = F.conv2d(x, sd[f"{prefix}.conv.weight"], sd[f"{prefix}.conv.bias"], stride=(4, 1), padding=(2, 0))
    else:
        le = x.shape[-1]
        if le % 4:
            x = F.pad(x, (0, 4 - le % 4))
        y = F.conv1d(x, sd[f"{prefix}.conv.weight"], sd[f"{prefix}.conv.bias"], stride=4, padding=2)
    y = F.gelu(y)
    _tap(taps, f"{prefix}.conv_gelu", y)
    if freq:
        B, C, Fr, T = y.shape
        y = y.permute(0, 2, 1, 3).reshape(-1, C, T)
    y = dconv(sd, f"{prefix}.dconv", y)
    if freq:
        y = y.view(B, Fr, C, T).permute(0, 2, 1, 3)
    _tap(taps, f"{prefix}.dconv_out", y)
    z = (F.conv2d if freq else F.conv1d)(y, sd[f"{prefix}.rewrite.weight"], sd[f"{prefix}.rewrite.bias"])
    return F.glu(z, dim=1)


def _mha(x_q, x_kv, w_in, b_in, w_out, b_out, heads: int):
    """nn.MultiheadAttention forward, batch_first, no mask, eval."""
    E = x_q.shape[-1]
    q = F.linear(x_q, w_in[:E], b_in[:E])
    k = F.linear(x_kv, w_in[E:2 * E], b_in[E:2 * E])
    v = F.linear(x_kv, w_in[2 * E:], b_in[2 * E:])
    B, Sq, _ = q.shape
    Sk = k.shape[1]
    hd = E // heads
    q = q.view(B, Sq, heads, hd).transpose(1, 2)
    k = k.view(B, Sk, heads, hd).transpose(1, 2)
    v = v.view(B, Sk, heads, hd).transpose(1, 2)
    att = torch.softmax((q @ k.transpose(-1, -2)) / math.sqrt(hd), dim=-1)
    o = (att @ v).transpose(1, 2).reshape(B, Sq, E)
    return F.linear(o, w_out, b_out)


def _ln(x, sd, p):
    return F.layer_norm(x, (x.shape[-1],), sd[f"{p}.weight"], sd[f"{p}.bias"], eps=1e-5)


def _my_group_norm(x, sd, p):
    """MyGroupNorm(1, C) on (B, T, C): statistics over all tokens x channels of a sample."""
    return F.group_norm(x.transpose(1, 2), 1, sd[f"{p}.weight"], sd[f"{p}.bias"], eps=1e-5).transpose(1, 2)


def _ffn(x, sd, p):
    return F.linear(F.gelu(F.linear(x, sd[f"{p}.linear1.weight"], sd[f"{p}.linear1.bias"])),
                    sd[f"{p}.linear2.weight"], sd[f"{p}.linear2.bias"])


def _self_layer(x, sd, p):
    h = _ln(x, sd, f"{p}.norm1")
    h = _mha(h, h, sd[f"{p}.self_attn.in_proj_weight"], sd[f"{p}.self_attn.in_proj_bias"],
             sd[f"{p}.self_attn.out_proj.weight"], sd[f"{p}.self_attn.out_proj.bias"], 8)
    x = x + sd[f"{p}.gamma_1.scale"] * h
    x = x + sd[f"{p}.gamma_2.scale"] * _ffn(_ln(x, sd, f"{p}.norm2"), sd, p)
    return _my_group_norm(x, sd, f"{p}.norm_out")


def _cross_layer(q, k, sd, p):
    kn = _ln(k, sd, f"{p}.norm2")
    h = _mha(_ln(q, sd, f"{p}.norm1"), kn, sd[f"{p}.cross_attn.in_proj_weight"],
             sd[f"{p}.cross_attn.in_proj_bias"], sd[f"{p}.cross_attn.out_proj.weight"],
             sd[f"{p}.cross_attn.out_proj.bias"], 8)
    x = q + sd[f"{p}.gamma_1.scale"] * h
    x = x + sd[f"{p}.gamma_2.scale"] * _ffn(_ln(x, sd, f"{p}.norm3"), sd, p)
    return _my_group_norm(x, sd, f"{p}.norm_out")


def cross_transformer(sd, x: torch.Tensor, xt: torch.Tensor, taps: Taps = None):
    """demucs CrossTransformerEncoder. x [B,512,Fr,T1], xt [B,512,T2]."""
    p = "htdemucs.crosstransformer"
    B, C, Fr, T1 = x.shape
    pe2 = create_2d_sin_embedding(C, Fr, T1).to(x).permute(0, 3, 2, 1).reshape(1, T1 * Fr, C)
    x = x.permute(0, 3, 2, 1).reshape(B, T1 * Fr, C)          # "b c fr t1 -> b (t1 fr) c"
    x = _ln(x, sd, f"{p}.norm_in") + pe2
    T2 = xt.shape[-1]
    xt = xt.permute(0, 2, 1)
    xt = _ln(xt, sd, f"{p}.norm_in_t") + create_sin_embedding(T2, C).to(xt).permute(1, 0, 2)
    _tap(taps, "xf_in", x)
    _tap(taps, "xf_in_t", xt)
    for idx in range(5):
        if idx % 2 == 0:
            x = _self_layer(x, sd, f"{p}.layers.{idx}")
            xt = _self_layer(xt, sd, f"{p}.layers_t.{idx}")
        else:
            old_x = x
            x = _cross_layer(x, xt, sd, f"{p}.layers.{idx}")
            xt = _cross_layer(xt, old_x, sd, f"{p}.layers_t.{idx}")
        _tap(taps, f"xf_layer{idx}", x)
        _tap(taps, f"xf_layer{idx}_t", xt)
    x = x.reshape(B, T1, Fr, C).permute(0, 3, 2, 1)
    xt = xt.permute(0, 2, 1)
    return x, xt


def encode(sd, x: torch.Tensor, xt: torch.Tensor, taps: Taps = None):
    """AudioTextHTDemucs._encode (ATHTDemucs_v2.py:190-236)."""
    saved, saved_t, lengths, lengths_t = [], [], [], []
    for idx in range(4):
        lengths.append(x.shape[-1])            # NB: the frame axis (ATHTDemucs_v2.py:198, quirk Q1)
        lengths_t.append(xt.shape[-1])
        xt = henc_layer(sd, f"htdemucs.tencoder.{idx}", xt, freq=False, taps=taps)
        saved_t.append(xt)
        x = henc_layer(sd, f"htdemucs.encoder.{idx}", x, freq=True, taps=taps)
        if idx == 0:
            frs = torch.arange(x.shape[-2])
            emb = (sd["htdemucs.freq_emb.embedding.weight"][frs] * 10.0).t()[None, :, :, None].expand_as(x)
            x = x + 0.2 * emb
        saved.append(x)
        _tap(taps, f"enc{idx}", x)
        _tap(taps, f"tenc{idx}", xt)
    b, c, f, t = x.shape
    x = F.conv1d(x.reshape(b, c, f * t), sd["htdemucs.channel_upsampler.weight"],
                 sd["htdemucs.channel_upsampler.bias"]).reshape(b, -1, f, t)
    xt = F.conv1d(xt, sd["htdemucs.channel_upsampler_t.weight"], sd["htdemucs.channel_upsampler_t.bias"])
    x, xt = cross_transformer(sd, x, xt, taps)
    x = F.conv1d(x.reshape(b, -1, f * t), sd["htdemucs.channel_downsampler.weight"],
                 sd["htdemucs.channel_downsampler.bias"]).reshape(b, -1, f, t)
    xt = F.conv1d(xt, sd["htdemucs.channel_downsampler_t.weight"], sd["htdemucs.channel_downsampler_t.bias"])
    _tap(taps, "x_enc", x)
    _tap(taps, "xt_enc", xt)
    return x, xt, saved, saved_t, lengths, lengths_t


# ----------------------------------------------------------------------------- text attention
def text_attend(sd, tokens: torch.Tensor, text_emb: torch.Tensor) -> torch.Tensor:
    """TextCrossAttention.forward_attend (ATHTDemucs_v2.py:38-48). tokens [B,S,384]."""
    p = "text_attn"
    q = _ln(tokens, sd, f"{p}.norm_q")
    if text_emb.dim() == 2:
        text_emb = text_emb.unsqueeze(1)
    k = F.linear(text_emb, sd[f"{p}.k_proj.weight"], sd[f"{p}.k_proj.bias"])
    v = F.linear(text_emb, sd[f"{p}.v_proj.weight"], sd[f"{p}.v_proj.bias"])
    qp = F.linear(q, sd[f"{p}.q_proj.weight"], sd[f"{p}.q_proj.bias"])
    E = qp.shape[-1]
    w_in, b_in = sd[f"{p}.attn.in_proj_weight"], sd[f"{p}.attn.in_proj_bias"]
    qq = F.linear(qp, w_in[:E], b_in[:E])
    kk = F.linear(k, w_in[E:2 * E], b_in[E:2 * E])
    vv = F.linear(v, w_in[2 * E:], b_in[2 * E:])
    B, S, _ = qq.shape
    Sk = kk.shape[1]
    H, hd = 8, E // 8
    qq = qq.view(B, S, H, hd).transpose(1, 2)
    kk = kk.view(B, Sk, H, hd).transpose(1, 2)
    vv = vv.view(B, Sk, H, hd).transpose(1, 2)
    att = torch.softmax((qq @ kk.transpose(-1, -2)) / math.sqrt(hd), dim=-1)
    o = (att @ vv).transpose(1, 2).reshape(B, S, E)
    attn_out = F.linear(o, sd[f"{p}.attn.out_proj.weight"], sd[f"{p}.attn.out_proj.bias"])
    out = tokens + attn_out
    mlp = F.linear(F.gelu(F.linear(out, sd[f"{p}.out_mlp.0.weight"], sd[f"{p}.out_mlp.0.bias"])),
                   sd[f"{p}.out_mlp.2.weight"], sd[f"{p}.out_mlp.2.bias"])
    out = out + mlp
    return _ln(out, sd, f"{p}.norm_out")


def text_attn(sd, x: torch.Tensor, xt: torch.Tensor, text_emb: torch.Tensor):
    """TextCrossAttention.forward (ATHTDemucs_v2.py:50-58)."""
    B, C, Fr, T = x.shape
    xs = x.permute(0, 2, 3, 1).reshape(B, Fr * T, C)            # b (f t) c
    xts = xt.permute(0, 2, 1)
    xs = text_attend(sd, xs, text_emb)
    xts = text_attend(sd, xts, text_emb)
    return xs.reshape(B, Fr, T, C).permute(0, 3, 1, 2), xts.permute(0, 2, 1)


# ----------------------------------------------------------------------------- decoders
def freq_decoder(sd, x, skips: List[torch.Tensor], target_lengths: List[int], taps: Taps = None):
    """FreqDecoder.forward (ATHTDemucs_v2.py:82-104)."""
    for i in range(4):
        p = f"freq_decoder.layers.{i}"
        x = F.conv_transpose2d(x, sd[f"{p}.0.weight"], sd[f"{p}.0.bias"], stride=(4, 1), padding=(2, 0))
        if i < 3:
            x = F.gelu(F.group_norm(x, 1, sd[f"{p}.1.weight"], sd[f"{p}.1.bias"], eps=1e-5))
        if i < len(target_lengths) and x.shape[2] != target_lengths[i]:
            x = F.interpolate(x, size=(target_lengths[i], x.shape[3]), mode="bilinear", align_corners=False)
        if i < len(skips):
            skip = skips[i]
            if skip.shape[1] != x.shape[1]:
                skip = skip[:, :x.shape[1]]
            if skip.shape[2:] != x.shape[2:]:
                skip = F.interpolate(skip, size=x.shape[2:], mode="bilinear", align_corners=False)
            x = x + skip * 0.1
        _tap(taps, f"fdec{i}", x)
    return x


def time_decoder(sd, x, skips: List[torch.Tensor], target_lengths: List[int], taps: Taps = None):
    """TimeDecoder.forward (ATHTDemucs_v2.py:125-139)."""
    for i in range(4):
        p = f"time_decoder.layers.{i}"
        x = F.conv_transpose1d(x, sd[f"{p}.0.weight"], sd[f"{p}.0.bias"], stride=4, padding=2)
        if i < 3:
            x = F.gelu(F.group_norm(x, 1, sd[f"{p}.1.weight"], sd[f"{p}.1.bias"], eps=1e-5))
        if i < len(target_lengths) and x.shape[2] != target_lengths[i]:
            x = F.interpolate(x, size=target_lengths[i], mode="linear", align_corners=False)
        if i < len(skips):
            skip = skips[i]
            if skip.shape[1] != x.shape[1]:
                skip = skip[:, :x.shape[1]]
            if skip.shape[2] != x.shape[2]:
                skip = F.interpolate(skip, size=x.shape[2], mode="linear", align_corners=False)
            x = x + skip * 0.1
        _tap(taps, f"tdec{i}", x)
    return x


# ----------------------------------------------------------------------------- forward
@torch.no_grad()
def forward(sd: Dict[str, torch.Tensor], wav: torch.Tensor, text_emb: torch.Tensor,
            taps: Taps = None) -> torch.Tensor:
    """AudioTextHTDemucs.forward (ATHTDemucs_v2.py:250-326) with the CLAP call (:282)
    replaced by the given (B,512) embedding."""
    original_length = wav.shape[-1]
    z = spec(wav)
    mag = magnitude(z)
    _tap(taps, "z", z)
    x = mag
    B, C, Fq, T_spec = x.shape
    mean = x.mean(dim=(1, 2, 3), keepdim=True)
    std = x.std(dim=(1, 2, 3), keepdim=True)
    x = (x - mean) / (1e-5 + std)
    xt = wav
    meant = xt.mean(dim=(1, 2), keepdim=True)
    stdt = xt.std(dim=(1, 2), keepdim=True)
    xt = (xt - meant) / (1e-5 + stdt)
    _tap(taps, "x_norm", x)
    _tap(taps, "xt_norm", xt)

    x_enc, xt_enc, saved, saved_t, lengths, lengths_t = encode(sd, x, xt, taps)
    x_cond, xt_cond = text_attn(sd, x_enc, xt_enc, text_emb)
    _tap(taps, "x_cond", x_cond)
    _tap(taps, "xt_cond", xt_cond)

    x_dec = freq_decoder(sd, x_cond, saved[::-1], lengths[::-1], taps)
    x_dec = F.conv2d(x_dec, sd["freq_out.weight"], sd["freq_out.bias"])
    _tap(taps, "x_dec", x_dec)
    x_dec = F.interpolate(x_dec, size=(Fq, T_spec), mode="bilinear", align_corners=False)
    mask = torch.sigmoid(x_dec)
    mag_stereo = mag[:, :2]
    masked_spec = mag_stereo * mask
    z_stereo = z[:, :2]
    phase = z_stereo / (mag_stereo + 1e-8)
    masked_z = masked_spec * phase
    _tap(taps, "masked_z", masked_z)
    freq_wav = ispec(masked_z, original_length)
    _tap(taps, "freq_wav", freq_wav)

    xt_dec = time_decoder(sd, xt_cond, saved_t[::-1], lengths_t[::-1], taps)
    xt_dec = F.conv1d(xt_dec, sd["time_out.weight"], sd["time_out.bias"])
    if xt_dec.shape[-1] != original_length:
        xt_dec = F.interpolate(xt_dec, size=original_length, mode="linear", align_corners=False)
    xt_dec = xt_dec * stdt + meant
    _tap(taps, "xt_dec", xt_dec)
    return freq_wav + xt_dec


def snr_db(est: torch.Tensor, ref: torch.Tensor) -> float:
    """Unclamped SNR, the formula of new_sdr_metric (/root/reference/src/loss.py:71-87)."""
    num = torch.sum(ref.double() ** 2)
    den = torch.sum((ref.double() - est.double()) ** 2)
    return float(10 * torch.log10((num + 1e-8) / (den + 1e-8)))
