"""GPU (-m gpu): out-of-bounds write guards (compute-sanitizer is closed on this pool): every output buffer of the stand-alone
C-ABI entry points sits between sentinel regions that must come back untouched, at sizes that are not multiples of the kernels'
tile / vector widths."""
import pytest
import torch

pytestmark = pytest.mark.gpu

import athtd_b200
from athtd_b200 import lib as alib, synthetic

GUARD = 4096
SENT = -12345.5


def _guarded(shape, dtype=torch.float32):
    n = 1
    for s in shape:
        n *= s
    buf = torch.full((n + 2 * GUARD,), SENT, dtype=dtype, device="cuda")
    return buf, buf[GUARD:GUARD + n].view(*shape)


def _intact(buf):
    return bool((buf[:GUARD] == SENT).all()) and bool((buf[-GUARD:] == SENT).all())


def _stream():
    return torch.cuda.current_stream().cuda_stream


@pytest.mark.parametrize("B,L", [(1, 4096), (2, 5001), (3, 40000), (1, 264600)])
def test_stft_and_fused_istft_stay_inside_their_buffers(B, L):
    eng = athtd_b200.Engine("cuda", "fp32")
    lib = alib.load()
    Tf = (L + 1023) // 1024
    wav = synthetic.make_inputs(9, B, L)[0].cuda()
    zb, Z = _guarded((B, Tf, 2048, 4))
    ob, out = _guarded((B, 2, L))
    stats = torch.zeros(2 * B, dtype=torch.float64, device="cuda")
    alib.check(lib.athtd_stft_cac(wav.data_ptr(), B, L, Z.data_ptr(), stats.data_ptr(), eng.tw.data_ptr(), eng.win.data_ptr(), _stream()))
    alib.check(lib.athtd_istft(Z.data_ptr(), B, L, 0, out.data_ptr(), eng.tw.data_ptr(), eng.win.data_ptr(), _stream()))
    torch.cuda.synchronize()
    assert _intact(zb) and _intact(ob)
    assert torch.isfinite(Z).all() and torch.isfinite(out).all() and (out != SENT).all()


@pytest.mark.parametrize("sr,C,T", [(48000, 1, 1001), (22050, 2, 4097), (96000, 1, 29999), (44100, 1, 777)])
def test_load_audio_stays_inside_its_buffer(sr, C, T):
    r = athtd_b200.DeviceResampler(sr, 44100)
    x = torch.randn(C, T, device="cuda")
    n_out = r.out_length(T)
    yb, y = _guarded((2, n_out))
    lib = alib.load()
    if r.identity:
        rc = lib.athtd_load_audio(x.data_ptr(), C, T, None, 1, 1, 0, 0, y.data_ptr(), 2, n_out, _stream())
    else:
        kt = r._kt_host.cuda()
        rc = lib.athtd_load_audio(x.data_ptr(), C, T, kt.data_ptr(), r.o, r.nw, r.taps, r.width, y.data_ptr(), 2, n_out, _stream())
    alib.check(rc)
    torch.cuda.synchronize()
    assert _intact(yb) and (y != SENT).all()


@pytest.mark.parametrize("T", [1, 70001, 500003])
def test_gather_and_overlap_add_stay_inside_their_buffers(T):
    plan = athtd_b200.segment_plan(T)
    tab = athtd_b200.OlaTables(plan, "cuda")
    n = len(plan.starts)
    mix = torch.randn(2, T, device="cuda")
    sb, segs = _guarded((n, 2, plan.chunk_len))
    lib = alib.load()
    alib.check(lib.athtd_gather_chunks(mix.data_ptr(), T, 2, tab.starts.data_ptr(), n, plan.chunk_len, segs.data_ptr(), _stream()))
    seg_out = torch.zeros(n + 1, 2, plan.chunk_len, device="cuda")
    seg_out[1:] = segs
    ob, out = _guarded((2, T))
    athtd_b200.chunk_ola(seg_out, 2 * plan.chunk_len, -1, tab, 0, T, out)
    torch.cuda.synchronize()
    assert _intact(sb) and _intact(ob)
    assert (segs != SENT).all() and (out != SENT).all()
    assert torch.allclose(out, mix, atol=1e-6)          # identity model: overlap-add of the chunks gives the track back


def test_forward_output_and_workspace_guards():
    """The plan's workspace and the output tensor between sentinels: a forward at a shape that is not a multiple of any tile."""
    m = athtd_b200.AudioTextHTDemucsB200(precision="bf16")
    m.load_state_dict(synthetic.make_state_dict(0), strict=False)
    m = m.cuda().eval()
    B, L, P = 3, 9001, 2
    wav = synthetic.make_inputs(3, B, L)[0].cuda()
    emb = torch.stack([synthetic.make_inputs(4 + p, B, 4096)[1] for p in range(P)], dim=1).cuda().contiguous()
    plan = m.engine().plan(B, L, P)
    ob, out = _guarded((B, P, 2, L))
    plan.forward(wav, emb, out)
    plan.forward(wav[:2].contiguous(), emb[:2].contiguous(), out[:2])
    torch.cuda.synchronize()
    assert _intact(ob) and torch.isfinite(out).all()
