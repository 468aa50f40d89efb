"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's evaluation metrics (torch, fp32), the checker for
audio-to-sheet-music_b200/metrics.py.  Pinned against the reference's own functions imported from
/root/reference/src/loss.py by oracle/make_metric_fixtures.py (fixtures: tests/golden/metrics.json)."""
import torch


def sdr_loss(estimated, target):
    """src/loss.py:9-30: -mean(clamp(10 log10((sum t^2 + 1e-8) / (sum (t - e)^2 + 1e-8)), -30, 30)) over items."""
    e = estimated.reshape(estimated.shape[0], -1)
    t = target.reshape(target.shape[0], -1)
    num = torch.sum(t ** 2, dim=-1)
    den = torch.sum((t - e) ** 2, dim=-1)
    return -torch.clamp(10 * torch.log10((num + 1e-8) / (den + 1e-8)), min=-30, max=30).mean()


def sisdr_loss(estimated, target):
    """src/loss.py:33-68: zero-mean both, project e on t, noise = orthogonal part, same clamp."""
    e = estimated.reshape(estimated.shape[0], -1)
    t = target.reshape(target.shape[0], -1)
    e = e - e.mean(dim=-1, keepdim=True)
    t = t - t.mean(dim=-1, keepdim=True)
    dot = torch.sum(e * t, dim=-1, keepdim=True)
    tt = torch.sum(t ** 2, dim=-1, keepdim=True)
    s_target = (dot / (tt + 1e-8)) * t
    e_noise = e - s_target
    v = 10 * torch.log10((torch.sum(s_target ** 2, dim=-1) + 1e-8) / (torch.sum(e_noise ** 2, dim=-1) + 1e-8))
    return -torch.clamp(v, min=-30, max=30).mean()


def new_sdr_metric(estimated, target):
    """src/loss.py:71-87: per-item SDR over (channels, time), unclamped."""
    num = torch.sum(target ** 2, dim=(1, 2))
    den = torch.sum((target - estimated) ** 2, dim=(1, 2))
    return 10 * torch.log10((num + 1e-8) / (den + 1e-8))


def make_case(seed: int, batch: int, channels: int, n: int, noise: float, gain: float = 1.0, offset: float = 0.0):
    """Seeded (estimate, target) pair: target = sinusoids + noise floor, estimate = gain * target + offset + noise."""
    g = torch.Generator().manual_seed(seed)
    t = torch.arange(n) / 44100.0
    tgt = 0.3 * torch.sin(6.2831853 * 220.0 * t) + 0.1 * torch.sin(6.2831853 * 1375.0 * t + 0.7)
    tgt = tgt.expand(batch, channels, n).clone() + 0.05 * torch.randn(batch, channels, n, generator=g)
    est = gain * tgt + offset + noise * torch.randn(batch, channels, n, generator=g)
    return est, tgt
