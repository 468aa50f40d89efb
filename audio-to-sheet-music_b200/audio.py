"""Host-side staging in front of the chunk loop: the reference's ``load_audio``
(/root/reference/app.py:113-126) = ``torchaudio.load`` -> ``torchaudio.transforms.Resample(sr, 44100)`` if the rates differ ->
``waveform.repeat(2, 1)`` if the file is mono.  File decoding stays the caller's (``torchaudio.load`` needs codecs); the
resampling and the channel repetition run on the device (``athtd_load_audio``: polyphase sinc FIR, one launch).

The filter bank is torchaudio's closed form (``torchaudio/functional/functional.py:_get_sinc_resample_kernel`` with its defaults:
sinc_interp_hann, lowpass_filter_width 6, rolloff 0.99), evaluated in float64 on the host like torchaudio does and rounded to
fp32; ``tests/test_host_logic.py`` checks it against torchaudio's own table bit for bit."""
from __future__ import annotations

import math
from typing import Dict, Tuple

import torch

from . import lib as _lib

TARGET_SR = 44100


def sinc_resample_kernel(orig_freq: int, new_freq: int, lowpass_filter_width: int = 6, rolloff: float = 0.99) -> Tuple[torch.Tensor, int, int, int]:
    """-> (kernel [new/g, 2*width + orig/g] fp32, width, orig/g, new/g) for the hann-windowed sinc interpolation."""
    if int(orig_freq) != orig_freq or int(new_freq) != new_freq or orig_freq < 1 or new_freq < 1:
        raise ValueError("sample rates must be positive integers")
    g = math.gcd(int(orig_freq), int(new_freq))
    o, nw = int(orig_freq) // g, int(new_freq) // g
    base = min(o, nw) * rolloff
    width = math.ceil(lowpass_filter_width * o / base)
    idx = torch.arange(-width, width + o, dtype=torch.float64)[None, None] / o
    t = torch.arange(0, -nw, -1)[:, None, None] / nw + idx      # int64 / int -> float32 phases, promoted to float64 (as torchaudio)
    t *= base
    t = t.clamp_(-lowpass_filter_width, lowpass_filter_width)
    window = torch.cos(t * math.pi / lowpass_filter_width / 2) ** 2
    t *= math.pi
    k = torch.where(t == 0, torch.tensor(1.0, dtype=torch.float64), t.sin() / t)
    k *= window * (base / o)
    return k.to(torch.float32)[:, 0].contiguous(), width, o, nw


class DeviceResampler:
    """``torchaudio.transforms.Resample(orig_freq, new_freq)`` (default arguments) + optional mono -> stereo on the device."""

    def __init__(self, orig_freq: int, new_freq: int = TARGET_SR):
        self.orig_freq, self.new_freq = int(orig_freq), int(new_freq)
        self.identity = self.orig_freq == self.new_freq
        if not self.identity:
            k, self.width, self.o, self.nw = sinc_resample_kernel(self.orig_freq, self.new_freq)
            self.taps = k.shape[1]
            self._kt_host = k.t().contiguous()              # [taps][new/g]: coalesced over the phases of a warp
        self._kt: Dict[str, torch.Tensor] = {}

    def out_length(self, T_in: int) -> int:
        return T_in if self.identity else -(-self.nw * T_in // self.o)

    def __call__(self, waveform: torch.Tensor, channels: int = None) -> torch.Tensor:
        """waveform [C, T] float32 on a CUDA device -> [channels or C, ceil(new * T / orig)]."""
        if not waveform.is_cuda:
            raise _lib.AthtdError("waveform must live on the CUDA device (no CPU fallback)")
        x = waveform.float().contiguous()
        C, T = x.shape
        Co = C if channels is None else channels
        y = torch.empty(Co, self.out_length(T), dtype=torch.float32, device=x.device)
        st = torch.cuda.current_stream(x.device).cuda_stream
        with torch.cuda.device(x.device):
            if self.identity:
                rc = _lib.load().athtd_load_audio(x.data_ptr(), C, T, None, 1, 1, 0, 0, y.data_ptr(), Co, y.shape[1], st)
            else:
                key = str(x.device)
                if key not in self._kt:
                    self._kt[key] = self._kt_host.to(x.device)
                rc = _lib.load().athtd_load_audio(x.data_ptr(), C, T, self._kt[key].data_ptr(), self.o, self.nw, self.taps, self.width,
                                                  y.data_ptr(), Co, y.shape[1], st)
        _lib.check(rc, "athtd_load_audio")
        return y


_RESAMPLERS: Dict[Tuple[int, int], DeviceResampler] = {}


def prepare_mixture(waveform: torch.Tensor, sample_rate: int, target_sr: int = TARGET_SR, device="cuda") -> Tuple[torch.Tensor, int]:
    """The tensor half of ``load_audio`` (app.py:113-126): (waveform [C, T] as torchaudio.load returns it, its rate) ->
    (stereo-or-wider [max(C, 2) if C == 1 else C, T'] float32 on the device at ``target_sr``, target_sr).  Like the reference,
    only a MONO file is widened (repeated to 2 channels); other channel counts pass through."""
    key = (int(sample_rate), int(target_sr))
    if key not in _RESAMPLERS:
        _RESAMPLERS[key] = DeviceResampler(*key)
    x = waveform.to(device, torch.float32)
    C = x.shape[0]
    return _RESAMPLERS[key](x, channels=2 if C == 1 else C), target_sr
