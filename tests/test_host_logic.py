"""CPU: host-side logic of the product (no CUDA compute): segment plan, ramp tables, module API /
state_dict layout, C-ABI exports."""
import os
import re

import pytest
import torch

import athtd_b200
from athtd_b200 import lib as alib
from oracle import ola, weights

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("T", [1, 66149, 66150, 198449, 198450, 198451, 264600, 264601, 10584000, 50803200, 158760000,
                               123457, 7000001])
def test_segment_plan_matches_oracle(T):
    mine = athtd_b200.segment_plan(T)
    ref = ola.chunk_plan(T)
    assert len(mine.starts) == len(ref)
    for k, c in enumerate(ref):
        assert (mine.starts[k], mine.ends[k], mine.actual_len[k], mine.fade_len[k]) == (c.start, c.end, c.actual_len, c.fade_len)
        assert mine.flags[k] == (1 if c.fade_in else 0) | (2 if c.fade_out else 0)


def test_segment_plan_app_overlap():
    mine = athtd_b200.segment_plan(1_000_000, 6.0, 0.1)     # app.py:29-33 / config.yaml:7
    ref = ola.chunk_plan(1_000_000, 6.0, 0.1)
    assert mine.starts == [c.start for c in ref] and mine.fade_len == [c.fade_len for c in ref]


def test_segment_plan_rejects_triple_overlap():
    with pytest.raises(ValueError):
        athtd_b200.segment_plan(10 ** 6, 6.0, 3.5)


def test_ramp_tables_bit_exact():
    plan = athtd_b200.segment_plan(240 * 44100)
    tab = athtd_b200.OlaTables(plan, "cpu")
    for k, c in enumerate(ola.chunk_plan(240 * 44100)):
        w = ola.chunk_weight(c)
        off = int(tab.ramp_off[k])
        if c.fade_in:
            assert torch.equal(tab.ramp_up[off:off + c.fade_len], w[:c.fade_len])
        if c.fade_out:
            assert torch.equal(tab.ramp_down[off:off + c.fade_len], w[-c.fade_len:])


def test_header_symbols_exported():
    hdr = open(os.path.join(ROOT, "include", "athtd.h")).read()
    declared = set(re.findall(r"\b(athtd_[a-z0-9_]+)\s*\(", hdr))
    bound = {s[0] for s in alib.SYMBOLS}
    assert declared == bound, declared ^ bound
    lib = alib.load()                      # dlopen + getattr on every symbol
    assert lib.athtd_version() >= 1
    assert lib.athtd_param_count() == 407


def test_param_table_is_the_reference_state_dict_layout():
    table = {n: numel for n, numel, _ in alib.param_table()}
    ref = {n: int(torch.tensor(s).prod()) for n, s, _ in weights.live_param_table()}
    assert table == ref


def test_module_state_dict_layout_and_dead_keys():
    model = athtd_b200.AudioTextHTDemucsB200()
    sd = weights.make_state_dict(0, include_dead=True)
    res = model.load_state_dict(sd, strict=False)
    assert res.missing_keys == []
    assert all(k.startswith(("clap.", "htdemucs.decoder.", "htdemucs.tdecoder.")) for k in res.unexpected_keys)
    live = {n for n, _, _ in weights.live_param_table()}
    assert set(model.state_dict().keys()) == live
    for k, v in model.state_dict().items():
        assert v.shape == sd[k].shape, k
    assert all(not p.requires_grad for p in model.htdemucs.parameters())
    assert any(p.requires_grad for p in model.text_attn.parameters())


def test_no_cpu_fallback():
    model = athtd_b200.AudioTextHTDemucsB200()
    with pytest.raises(athtd_b200.AthtdError):
        model(torch.zeros(1, 2, 44100), torch.zeros(1, 512))
    assert isinstance(athtd_b200.B200SeparationModel.__mro__[1], type) and issubclass(athtd_b200.B200SeparationModel, athtd_b200.SeparationModel)


def test_workspace_query():
    lib = alib.load()
    assert lib.athtd_workspace_bytes(1, 264600, 1, 0) > 0
    assert lib.athtd_workspace_bytes(1, 100, 1, 0) == -1          # too short: explicit error, not UB
    assert b"4096" in lib.athtd_last_error()


def test_positional_tables_match_oracle():
    from athtd_b200.engine import sin_embedding_1d, sin_embedding_2d
    from oracle.demucs_shim import create_2d_sin_embedding, create_sin_embedding
    ref2 = create_2d_sin_embedding(512, 8, 40).permute(0, 3, 2, 1).reshape(320, 512)
    assert torch.equal(sin_embedding_2d(512, 8, 40), ref2)
    ref1 = create_sin_embedding(157, 512).permute(1, 0, 2)[0]
    assert torch.equal(sin_embedding_1d(157, 512), ref1)


def test_fade_mask_matches_torchaudio_fade():
    """oracle/ola.fade_mask restates torchaudio.transforms.Fade(..., "linear") (the transform test_inference.py:126-134
    applies); pinned against the real transform when torchaudio is importable."""
    torchaudio = pytest.importorskip("torchaudio")
    from torchaudio.transforms import Fade
    from oracle import ola
    for (L, a, b) in [(1000, 0, 100), (1000, 100, 0), (5000, 441, 441), (300, 100, 250), (264600, 4410, 4410)]:
        x = torch.randn(2, L)
        assert torch.equal(Fade(a, b, "linear")(x), ola.fade_mask(L, a, b) * x)


def test_fade_inference_identity_model_reconstructs_interior():
    """With an identity model the faded chunks of test_inference.py sum to the input wherever fade-in + fade-out = 1
    up to fp32 rounding (the loop has no weight normalisation)."""
    from oracle import ola
    T = 44100 * 3 + 5000
    x = torch.randn(2, T)
    y = ola.fade_inference(lambda c: c, x, 1.0, 0.1)
    assert (y - x).abs().max() < 1e-5


def test_bench_reference_arm_prints_one_json_line():
    """bench.py --impl reference (the reference algorithm on the host cores, tier contract 4): stdout is exactly one JSON
    line carrying the keys the driver reads."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--seconds", "12"],
                       capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "x realtime" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


def test_separation_model_abc_keeps_the_reference_contract():
    """benchmark.py:81-115: abstract separate / separate_all / name; a reference-style plugin that implements exactly those three
    instantiates, one that forgets ``separate`` does not."""
    import athtd_b200

    class Plugin(athtd_b200.SeparationModel):
        def separate(self, mixture, stem_name):
            return mixture

        def separate_all(self, mixture):
            return {s: mixture for s in athtd_b200.STEMS}

        @property
        def name(self):
            return "plugin"

    assert Plugin().name == "plugin"
    assert athtd_b200.SeparationModel.__abstractmethods__ == frozenset({"separate", "separate_all", "name"})

    class Broken(athtd_b200.SeparationModel):
        def separate_all(self, mixture):
            return {}

        @property
        def name(self):
            return "broken"

    with pytest.raises(TypeError):
        Broken()
    assert not hasattr(athtd_b200.SeparationModel, "separate_fade")
    for m in ("separate", "separate_all", "separate_fade", "separate_span", "separate_span_host"):
        assert callable(getattr(athtd_b200.B200SeparationModel, m))


def test_product_package_never_imports_the_oracle():
    """The product path (package + bench.py's GPU arm) must not route through oracle/: only tests, smoke() and the CPU legs of
    bench.py may."""
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "audio-to-sheet-music_b200")
    pat = re.compile(r"^\s*(from|import)\s+oracle\b", re.M)
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            assert not pat.search(open(os.path.join(pkg, fn)).read()), fn
    bench = open(os.path.join(root, "bench.py")).read()
    for m in pat.finditer(bench):
        fn_start = bench.rfind("\ndef ", 0, m.start())
        name = bench[fn_start:bench.find("(", fn_start)].split()[-1]
        assert name in ("cpu_reference_rate", "cpu_stft_rate"), name


def test_bench_track_parts_tile_the_track():
    from athtd_b200 import synthetic
    a = synthetic.make_track(1.0)
    assert a.shape == (2, 44100) and torch.equal(a, synthetic.make_track_part(0, 44100))
    b = synthetic.make_track_part(1000, 3000)
    tone = (a - 0.0)[:, 1000:3000] - b           # same tone phase, different noise draw
    assert float(tone.abs().max()) < 1.0 and b.shape == (2, 2000)


@pytest.mark.parametrize("orig,new", [(48000, 44100), (22050, 44100), (32000, 44100), (96000, 44100), (8000, 44100), (44100, 16000)])
def test_resample_filter_bank_is_torchaudios(orig, new):
    """load_audio (app.py:113-126) resamples with torchaudio.transforms.Resample defaults: the device kernel's filter bank must be
    the table torchaudio itself builds, bit for bit."""
    import math
    from torchaudio.functional import functional as Fn
    g = math.gcd(orig, new)
    ref, width = Fn._get_sinc_resample_kernel(orig, new, g)
    k, w, o, nw = athtd_b200.audio.sinc_resample_kernel(orig, new)
    assert (w, o, nw) == (width, orig // g, new // g)
    assert torch.equal(ref[:, 0], k)
    r = athtd_b200.DeviceResampler(orig, new)
    assert r.out_length(1000) == math.ceil(new * 1000 / orig) and r.taps == 2 * width + orig // g
