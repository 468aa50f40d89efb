"""GPU (-m gpu): the tensor half of the reference's load_audio (app.py:113-126) -- resample to 44.1 kHz + mono -> stereo -- on the
device, against torchaudio.transforms.Resample on the CPU (the library the reference calls) as the oracle."""
import pytest
import torch
import torchaudio

pytestmark = pytest.mark.gpu

import athtd_b200


@pytest.mark.parametrize("sr,C,T", [(48000, 2, 200001), (22050, 1, 50000), (32000, 2, 77777), (96000, 1, 123456), (8000, 2, 4000),
                                    (44100, 1, 30000), (44100, 2, 30000), (48000, 1, 37)])
def test_prepare_mixture_matches_torchaudio_resample(sr, C, T):
    g = torch.Generator().manual_seed(sr + C + T)
    wav = torch.randn(C, T, generator=g).clamp_(-1, 1)
    ref = wav
    if sr != 44100:
        ref = torchaudio.transforms.Resample(sr, 44100)(wav)                # app.py:118-120
    if ref.shape[0] == 1:
        ref = ref.repeat(2, 1)                                             # app.py:123-124
    out, out_sr = athtd_b200.prepare_mixture(wav, sr, device="cuda")
    assert out_sr == 44100 and out.shape == ref.shape and out.is_cuda
    if sr == 44100:
        assert torch.equal(out.cpu(), ref)
    else:
        assert (out.cpu() - ref).abs().max() < 2e-5                        # fp32 FIR, different summation order than conv1d
