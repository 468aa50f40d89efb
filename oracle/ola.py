"""TEST INFRASTRUCTURE (oracle) -- restatement of the reference's track-level chunk loop.

Follows OurModel._chunked_inference (/root/reference/benchmark.py:155-204; the same
algorithm at /root/reference/app.py:129-178 with overlap 0.1 s): 6 s chunks, stride
chunk_len - overlap, zero-padded tail, linear fade weights from ``torch.linspace``,
weighted overlap-add and division by the clamped accumulated weight.  ``separate_all``
follows benchmark.py:210-215 (one full pass per stem).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this module.
"""
from __future__ import annotations

from typing import Callable, Dict, List, NamedTuple

import torch
import torch.nn.functional as F

SAMPLE_RATE = 44100
STEMS = ["drums", "bass", "other", "vocals"]     # benchmark.py:58


class Chunk(NamedTuple):
    start: int
    end: int
    actual_len: int
    fade_len: int
    fade_in: bool
    fade_out: bool


def chunk_plan(T: int, segment_seconds: float = 6.0, overlap_seconds: float = 1.5,
               sample_rate: int = SAMPLE_RATE) -> List[Chunk]:
    """Index arithmetic of benchmark.py:158-198."""
    chunk_len = int(sample_rate * segment_seconds)
    overlap_frames = int(overlap_seconds * sample_rate)
    plan = []
    start = 0
    while start < T:
        end = min(start + chunk_len, T)
        actual_len = end - start
        fade_len = min(overlap_frames, actual_len // 2)
        plan.append(Chunk(start, end, actual_len, fade_len,
                          bool(start > 0 and fade_len > 0), bool(end < T and fade_len > 0)))
        start += chunk_len - overlap_frames
    return plan


def chunk_weight(c: Chunk) -> torch.Tensor:
    """benchmark.py:185-192."""
    w = torch.ones(c.actual_len)
    if c.fade_in:
        w[:c.fade_len] = torch.linspace(0, 1, c.fade_len)
    if c.fade_out:
        w[-c.fade_len:] = torch.linspace(1, 0, c.fade_len)
    return w


def chunked_inference(model_fn: Callable[[torch.Tensor], torch.Tensor], mixture: torch.Tensor,
                      segment_seconds: float = 6.0, overlap_seconds: float = 1.5,
                      sample_rate: int = SAMPLE_RATE) -> torch.Tensor:
    """benchmark.py:155-204 with ``model_fn(chunk[1,C,chunk_len]) -> [1,C,chunk_len]``."""
    C, T = mixture.shape
    chunk_len = int(sample_rate * segment_seconds)
    output = torch.zeros(C, T)
    weight = torch.zeros(T)
    for c in chunk_plan(T, segment_seconds, overlap_seconds, sample_rate):
        chunk = mixture[:, c.start:c.end].unsqueeze(0)
        if chunk.shape[-1] < chunk_len:
            chunk = F.pad(chunk, (0, chunk_len - chunk.shape[-1]))
        out = model_fn(chunk).squeeze(0)[:, :c.actual_len]
        w = chunk_weight(c)
        output[:, c.start:c.end] += out * w
        weight[c.start:c.end] += w
    weight = weight.clamp(min=1e-8)
    return output / weight


def fade_mask(length: int, fade_in_len: int, fade_out_len: int) -> torch.Tensor:
    """torchaudio.transforms.Fade(fade_in_len, fade_out_len, "linear") applied to ones(length):
    cat(linspace(0,1,n_in), ones).clamp(0,1) * cat(ones, -linspace(0,1,n_out)+1).clamp(0,1)
    (torchaudio/transforms/_transforms.py Fade._fade_in / _fade_out / forward)."""
    fin = torch.cat((torch.linspace(0, 1, fade_in_len), torch.ones(length - fade_in_len))).clamp_(0, 1)
    fout = torch.cat((torch.ones(length - fade_out_len), -torch.linspace(0, 1, fade_out_len) + 1)).clamp_(0, 1)
    return fin * fout


def fade_inference(model_fn: Callable[[torch.Tensor], torch.Tensor], mixture: torch.Tensor,
                   segment_seconds: float = 6.0, overlap_seconds: float = 0.1, sample_rate: int = SAMPLE_RATE) -> torch.Tensor:
    """test_inference.py:96-141 for one stem: no zero padding of the last chunk (``model_fn`` sees its true length),
    Fade(fade_in = overlap iff start > 0, fade_out = overlap iff end < T), ``final[:, start:end] += out``, no normalisation."""
    C, length = mixture.shape
    chunk_len = int(sample_rate * segment_seconds)
    overlap_frames = int(overlap_seconds * sample_rate)
    final = torch.zeros(C, length)
    start = 0
    while start < length:
        end = min(start + chunk_len, length)
        chunk = mixture[:, start:end].unsqueeze(0)
        out = model_fn(chunk)
        fade_in = 0 if start == 0 else overlap_frames
        fade_out = overlap_frames if end < length else 0
        out = fade_mask(out.shape[-1], fade_in, fade_out) * out
        final[:, start:end] += out.squeeze(0)
        start += chunk_len - overlap_frames
    return final


def separate_all(model_fn_for: Callable[[str], Callable], mixture: torch.Tensor) -> Dict[str, torch.Tensor]:
    """benchmark.py:210-215: one complete chunked pass per stem prompt."""
    return {stem: chunked_inference(model_fn_for(stem), mixture) for stem in STEMS}
