"""Seeded synthetic ``state_dict`` and inputs for the live hot path (benchmarks, smoke run, tests).

No pretrained weights are reachable (no network; the checkpoint blob is missing from the
reference, /root/reference/.MISSING_LARGE_BLOBS), so the bench, the smoke run and the parity tests
use this recipe: PyTorch-default-like uniform inits, demucs
``rescale_module(0.1)`` on the HTDemucs convolutions, then every LayerScale raised to U(0.05, 0.5)
and every norm affine perturbed so that each branch contributes measurably.
Key names / shapes are those of the reference module
(/root/reference/src/models/stem_separation/AudioTextHTDemucs_Full.txt:3-629,
 ATHTDemucs_v2.py:178-188).  Pure torch-on-CPU data generation: nothing here computes the path.
"""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import torch

CH = [48, 96, 192, 384]


def live_param_table() -> List[Tuple[str, Tuple[int, ...], str]]:
    """(name, shape, kind) for the 407 live tensors; kind in
    {conv_w, conv_b, lin_w, lin_b, norm_w, norm_b, scale, emb}."""
    t: List[Tuple[str, Tuple[int, ...], str]] = []

    def conv(name, shape):
        t.append((f"{name}.weight", shape, "conv_w"))
        t.append((f"{name}.bias", (shape[0],), "conv_b"))

    def convtr(name, shape):
        t.append((f"{name}.weight", shape, "convtr_w"))
        t.append((f"{name}.bias", (shape[1],), "conv_b"))

    def lin(name, o, i):
        t.append((f"{name}.weight", (o, i), "lin_w"))
        t.append((f"{name}.bias", (o,), "lin_b"))

    def norm(name, c):
        t.append((f"{name}.weight", (c,), "norm_w"))
        t.append((f"{name}.bias", (c,), "norm_b"))

    for branch, freq in (("encoder", True), ("tencoder", False)):
        for i in range(4):
            cin = (4 if freq else 2) if i == 0 else CH[i - 1]
            c = CH[i]
            p = f"htdemucs.{branch}.{i}"
            conv(f"{p}.conv", (c, cin, 8, 1) if freq else (c, cin, 8))
            conv(f"{p}.rewrite", (2 * c, c, 1, 1) if freq else (2 * c, c, 1))
            for d in range(2):
                q = f"{p}.dconv.layers.{d}"
                conv(f"{q}.0", (c // 8, c, 3))
                norm(f"{q}.1", c // 8)
                conv(f"{q}.3", (2 * c, c // 8, 1))
                norm(f"{q}.4", 2 * c)
                t.append((f"{q}.6.scale", (c,), "scale"))
    t.append(("htdemucs.freq_emb.embedding.weight", (512, 48), "emb"))
    conv("htdemucs.channel_upsampler", (512, 384, 1))
    conv("htdemucs.channel_downsampler", (384, 512, 1))
    conv("htdemucs.channel_upsampler_t", (512, 384, 1))
    conv("htdemucs.channel_downsampler_t", (384, 512, 1))
    x = "htdemucs.crosstransformer"
    norm(f"{x}.norm_in", 512)
    norm(f"{x}.norm_in_t", 512)
    for stack in ("layers", "layers_t"):
        for i in range(5):
            p = f"{x}.{stack}.{i}"
            a = "self_attn" if i % 2 == 0 else "cross_attn"
            t.append((f"{p}.{a}.in_proj_weight", (1536, 512), "lin_w"))
            t.append((f"{p}.{a}.in_proj_bias", (1536,), "lin_b"))
            lin(f"{p}.{a}.out_proj", 512, 512)
            lin(f"{p}.linear1", 2048, 512)
            lin(f"{p}.linear2", 512, 2048)
            norm(f"{p}.norm1", 512)
            norm(f"{p}.norm2", 512)
            if i % 2 == 1:
                norm(f"{p}.norm3", 512)
            norm(f"{p}.norm_out", 512)
            t.append((f"{p}.gamma_1.scale", (512,), "scale"))
            t.append((f"{p}.gamma_2.scale", (512,), "scale"))
    lin("text_attn.q_proj", 384, 384)
    lin("text_attn.k_proj", 384, 512)
    lin("text_attn.v_proj", 384, 512)
    t.append(("text_attn.attn.in_proj_weight", (1152, 384), "lin_w"))
    t.append(("text_attn.attn.in_proj_bias", (1152,), "lin_b"))
    lin("text_attn.attn.out_proj", 384, 384)
    lin("text_attn.out_mlp.0", 384, 384)
    lin("text_attn.out_mlp.2", 384, 384)
    norm("text_attn.norm_q", 384)
    norm("text_attn.norm_out", 384)
    dch = [384, 192, 96, 48, 4]
    for i in range(4):
        convtr(f"freq_decoder.layers.{i}.0", (dch[i], dch[i + 1], 8, 1))
        if i < 3:
            norm(f"freq_decoder.layers.{i}.1", dch[i + 1])
    for i in range(4):
        convtr(f"time_decoder.layers.{i}.0", (dch[i], dch[i + 1], 8))
        if i < 3:
            norm(f"time_decoder.layers.{i}.1", dch[i + 1])
    conv("freq_out", (2, 4, 1, 1))
    conv("time_out", (2, 4, 1))
    return t


def make_state_dict(seed: int = 0, include_dead: bool = False) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}

    def uni(shape, b):
        return (torch.rand(shape, generator=g) * 2 - 1) * b

    for name, shape, kind in live_param_table():
        if kind in ("conv_w", "lin_w"):
            fan_in = int(torch.tensor(shape[1:]).prod())
            w = uni(shape, 1.0 / math.sqrt(fan_in))
        elif kind == "convtr_w":
            fan_in = int(torch.tensor((shape[1],) + tuple(shape[2:])).prod())
            w = uni(shape, 1.0 / math.sqrt(fan_in))
        elif kind in ("conv_b", "lin_b"):
            w = uni(shape, 0.05)
        elif kind == "norm_w":
            w = 0.5 + torch.rand(shape, generator=g)
        elif kind == "norm_b":
            w = 0.1 * torch.randn(shape, generator=g)
        elif kind == "scale":
            w = 0.05 + 0.45 * torch.rand(shape, generator=g)
        elif kind == "emb":
            w = torch.cumsum(torch.randn(shape, generator=g), dim=0)
            w = w / torch.arange(1, shape[0] + 1).float().sqrt()[:, None] / 10.0
        else:
            raise AssertionError(kind)
        sd[name] = w.float().contiguous()
    # demucs rescale_module(0.1) on the HTDemucs convolutions
    for name in list(sd):
        if name.startswith("htdemucs.") and name.endswith(".weight") and sd[name].dim() >= 3:
            s = (sd[name].std() / 0.1) ** 0.5
            sd[name] = sd[name] / s
            sd[name[:-6] + "bias"] = sd[name[:-6] + "bias"] / s
    if include_dead:  # keys a real checkpoint carries but the path never reads
        sd["htdemucs.decoder.0.conv_tr.weight"] = torch.zeros(384, 192, 8, 1)
        sd["htdemucs.tdecoder.0.conv_tr.weight"] = torch.zeros(384, 192, 8)
        sd["clap.text_model.embeddings.word_embeddings.weight"] = torch.zeros(8, 8)
    return sd


def make_inputs(seed: int, batch: int, length: int, emb_norm: bool = True):
    """Synthetic stereo 44.1 kHz mixture (Gaussian sigma 0.1 + sinusoids) and 512-d embeddings
    (SURVEY.md section 8d "value distributions")."""
    g = torch.Generator().manual_seed(seed)
    wav = 0.1 * torch.randn(batch, 2, length, generator=g)
    t = torch.arange(length).float() / 44100.0
    for k, f0 in enumerate((110.0, 440.0, 1760.0, 5000.0)):
        amp = 0.2 / (k + 1)
        ph = torch.rand(batch, 2, 1, generator=g) * 6.2831853
        wav = wav + amp * torch.sin(6.2831853 * f0 * t[None, None, :] + ph)
    emb = torch.randn(batch, 512, generator=g)
    if emb_norm:
        emb = emb / emb.norm(dim=-1, keepdim=True)
    return wav.float().contiguous(), emb.float().contiguous()


def state_dict_checksum(sd: Dict[str, torch.Tensor]) -> float:
    return float(sum(float(v.double().abs().sum()) for k, v in sorted(sd.items())))


def make_track_part(lo: int, hi: int, seed: int = 4, sample_rate: int = 44100) -> torch.Tensor:
    """Samples [lo, hi) of the bench's synthetic stereo track: Gaussian sigma 0.1 (generator seeded per part, so a rank can
    produce its own span without the whole track) plus a 220 Hz tone with global phase."""
    g = torch.Generator().manual_seed(seed + 1000003 * lo)
    part = 0.1 * torch.randn(2, hi - lo, generator=g)
    part += 0.2 * torch.sin(6.2831853 * 220.0 * (torch.arange(lo, hi, dtype=torch.float64) / sample_rate)).float()
    return part.float().contiguous()


def make_track(seconds: float, seed: int = 4, sample_rate: int = 44100) -> torch.Tensor:
    """The bench's synthetic stereo track [2, T] (bench.py, BASELINE configs 3-5) as one part."""
    return make_track_part(0, int(seconds * sample_rate), seed, sample_rate)


def make_prompt_embeddings(prompts: int, seed: int = 2, emb_norm: bool = True) -> torch.Tensor:
    """[P, 512] synthetic text embeddings (the bench's stand-in for the CLAP outputs of P prompts)."""
    _, emb = make_inputs(seed, prompts, 4096, emb_norm=emb_norm)
    return emb
