"""BASELINE.json configs[1]: STFT -> iSTFT round-trip microbench, nfft 4096 / hop 1024, 64 stereo 6 s segments on one B200.

    python tools/stft_bench.py [--B 64] [--iters 20]

Times athtd_stft_cac (htdemucs._spec + _magnitude) and athtd_istft (htdemucs._ispec) through the C ABI with CUDA events
on the launching stream and reports achieved HBM GB/s against the ALGORITHMIC bytes of SURVEY.md 8(d):
read wav + write Z + read Z + write wav = 2 x (B*2*L*4) + 2 x (B*Tf*2048*4*4) bytes (1.357 GB at B = 64).
The L2 (126 MB) is flushed between iterations by the kernels themselves: every buffer of the round trip is larger than L2
(wav 135 MB, Z 543 MB).  Correctness of each direction against the oracle is tests/test_gpu_kernels.py; here the output is
only checked against torch.stft / torch.istft evaluated on the GPU with the reference's padding rules (max-abs printed).
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import athtd_b200
from athtd_b200 import lib as alib


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=64)
    ap.add_argument("--L", type=int, default=264600)
    ap.add_argument("--iters", type=int, default=20)
    a = ap.parse_args()
    B, L = a.B, a.L
    Tf = (L + 1023) // 1024
    lib = alib.load()
    eng = athtd_b200.Engine("cuda", "fp32")
    g = torch.Generator(device="cuda").manual_seed(3)
    wav = torch.randn(B, 2, L, device="cuda", generator=g)
    Z = torch.empty(B, Tf, 2048, 4, device="cuda")
    frames = torch.empty(B * 2 * Tf * 4096, device="cuda")
    out = torch.empty(B, 2, L, device="cuda")
    stats = torch.zeros(2 * B, dtype=torch.float64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream

    def fwd():
        alib.check(lib.athtd_stft_cac(wav.data_ptr(), B, L, Z.data_ptr(), stats.data_ptr(), eng.tw.data_ptr(), eng.win.data_ptr(), st))

    def inv():
        alib.check(lib.athtd_istft(Z.data_ptr(), B, L, frames.data_ptr(), out.data_ptr(), eng.tw.data_ptr(), eng.win.data_ptr(), st))

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / a.iters

    ms_f = timed(fwd)
    ms_i = timed(inv)
    ms_rt = timed(lambda: (fwd(), inv()))
    wav_b = B * 2 * L * 4
    z_b = B * Tf * 2048 * 4 * 4
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = peaks.get("hbm_gbs", 6549.8)
    # reference semantics on the GPU (demucs spec.py / htdemucs._spec): only a sanity number, parity lives in tests/
    x = wav[:4]
    hl = 1024
    le = (L + hl - 1) // hl
    pad = hl // 2 * 3
    xp = torch.nn.functional.pad(x, (pad, pad + le * hl - L), mode="reflect")
    z = torch.stft(xp.reshape(-1, xp.shape[-1]), 4096, hl, window=torch.hann_window(4096, device="cuda"), win_length=4096, normalized=True,
                   center=True, return_complex=True, pad_mode="reflect").view(4, 2, 2049, -1)[:, :, :-1, 2:2 + le]
    zr = torch.view_as_real(z)
    ref = torch.stack([zr[:, 0, :, :, 0], zr[:, 0, :, :, 1], zr[:, 1, :, :, 0], zr[:, 1, :, :, 1]], dim=-1).permute(0, 2, 1, 3)
    err = float((Z[:4] - ref).abs().max())
    res = {
        "workload": f"STFT->iSTFT round trip, nfft 4096 hop 1024, {B} stereo segments of {L} samples",
        "stft_ms": ms_f, "istft_ms": ms_i, "round_trip_ms": ms_rt,
        "stft_GBs": (wav_b + z_b) / ms_f / 1e6, "istft_GBs": (wav_b + z_b) / ms_i / 1e6,
        "round_trip_GBs": 2 * (wav_b + z_b) / ms_rt / 1e6,
        "algorithmic_bytes": 2 * (wav_b + z_b), "hbm_peak_GBs": peak,
        "frac_of_peak": 2 * (wav_b + z_b) / ms_rt / 1e6 / peak,
        "istft_scratch_bytes_not_counted": int(frames.numel() * 4 * 2),
        "stft_max_abs_vs_torch_stft": err,
    }
    print(json.dumps(res))


if __name__ == "__main__":
    main()
