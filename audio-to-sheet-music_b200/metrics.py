"""Evaluation metrics of the reference (src/loss.py:9-87, wrappers benchmark.py:555-588) on the device.

Same names, argument meaning and return values as the reference functions; the waveforms stay on the GPU and one
bandwidth-bound kernel (``athtd_sdr_sums``) produces the six fp64 sums every metric is a closed form of:

    sdr_loss        -mean(clamp(10 log10((sum t^2 + d) / (sum (t-e)^2 + d)), -30, 30))          src/loss.py:9-30
    sisdr_loss      zero-mean projection of e on t, same clamp                                   src/loss.py:33-68
    new_sdr_metric  unclamped per-item SDR over (channels, time)                                 src/loss.py:71-87

The reference evaluates the SI-SDR element-wise in fp32; here the projection is expanded into moments
(sum e't' = sum e t - n mu_e mu_t, ...) in fp64, which agrees to < 1e-3 dB away from the +-30 dB clamp.
"""
from __future__ import annotations

import torch

from . import lib as _lib

DELTA = 1e-8


def _sums(estimated: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    if estimated.shape != target.shape:
        raise ValueError("estimated and target must have the same shape")
    if not (estimated.is_cuda and target.is_cuda):
        raise _lib.AthtdError("athtd_b200.metrics runs on CUDA tensors only (no CPU fallback)")
    items = estimated.shape[0]
    e = estimated.reshape(items, -1).float().contiguous()
    t = target.reshape(items, -1).float().contiguous()
    sums = torch.empty(items, 6, dtype=torch.float64, device=e.device)
    st = torch.cuda.current_stream(e.device).cuda_stream
    _lib.check(_lib.load().athtd_sdr_sums(e.data_ptr(), t.data_ptr(), items, e.shape[1], sums.data_ptr(), st), "athtd_sdr_sums")
    return sums, e.shape[1]


def _sdr_db(sums: torch.Tensor) -> torch.Tensor:
    return 10.0 * torch.log10((sums[:, 2] + DELTA) / (sums[:, 5] + DELTA))


def new_sdr_metric(estimated: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """(batch, channels, time) x2 -> per-item SDR in dB (batch,), unclamped (src/loss.py:71-87)."""
    sums, _ = _sums(estimated, target)
    return _sdr_db(sums).float()


def sdr_loss(estimated: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """Negative mean SDR, clamped to [-30, 30] dB per item (src/loss.py:9-30)."""
    sums, _ = _sums(estimated, target)
    return -torch.clamp(_sdr_db(sums), min=-30, max=30).mean().float()


def sisdr_loss(estimated: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """Negative mean SI-SDR, clamped to [-30, 30] dB per item (src/loss.py:33-68)."""
    sums, n = _sums(estimated, target)
    st, se, stt, see, set_, _ = (sums[:, i] for i in range(6))
    mt, me = st / n, se / n
    dot = set_ - n * me * mt                      # sum e' t'
    tt = stt - n * mt * mt                        # sum t'^2
    ee = see - n * me * me                        # sum e'^2
    alpha = dot / (tt + DELTA)
    s_target = alpha * alpha * tt
    e_noise = torch.clamp(ee - 2.0 * alpha * dot + alpha * alpha * tt, min=0.0)
    sisdr = 10.0 * torch.log10((s_target + DELTA) / (e_noise + DELTA))
    return -torch.clamp(sisdr, min=-30, max=30).mean().float()


def compute_sdr(estimate: torch.Tensor, reference: torch.Tensor) -> float:
    """(C, T) x2 -> SDR in dB (benchmark.py:555-570)."""
    return -sdr_loss(estimate.unsqueeze(0), reference.unsqueeze(0)).item()


def compute_sisdr(estimate: torch.Tensor, reference: torch.Tensor) -> float:
    """(C, T) x2 -> SI-SDR in dB (benchmark.py:573-588)."""
    return -sisdr_loss(estimate.unsqueeze(0), reference.unsqueeze(0)).item()
