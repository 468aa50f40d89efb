"""CPU: pin ``oracle/demucs_shim.py`` (the restated ``demucs==4.0.1`` half of the oracle, 83 % of the path's FLOPs)
to INDEPENDENT implementations that ship in this image, since the demucs source itself is absent:

* torchaudio's port of the same layers (``torchaudio/models/_hdemucs.py``: ``_HEncLayer``, ``_DConv``, ``_ScaledEmbedding``,
  ``_rescale_module``, ``_spectro`` / ``_ispectro``, ``HDemucs._spec`` / ``_ispec`` / ``_magnitude``) -- bit-equal;
* ``torch.nn.TransformerEncoderLayer(norm_first=True)`` + an explicit ``GroupNorm(1, 512)`` over (tokens x channels) with the
  LayerScale folded into the projection weights -- an independent composition of ``MyTransformerEncoderLayer``;
* ``F.scaled_dot_product_attention`` + ``F.layer_norm`` for ``CrossTransformerEncoderLayer``;
* the parameter totals the reference's own torchinfo dumps print
  (/root/reference/src/models/stem_separation/HTDemucs_Fwd_Pass.txt:69,73,147);
* closed forms of the sinusoidal embeddings evaluated in float64 with numpy.
The functional oracle (``oracle/athtd_oracle.py``, manual softmax attention) is then checked against the shim modules."""
import math
from types import SimpleNamespace

import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F
from torchaudio.models import _hdemucs as ta

from oracle import athtd_oracle, demucs_shim as shim


def _seed_module(mod: nn.Module, seed: int) -> None:
    """Random parameters with every branch alive (LayerScale raised, norm affines perturbed)."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in mod.named_parameters():
            if name.endswith("scale"):
                p.copy_(0.05 + 0.45 * torch.rand(p.shape, generator=g))
            elif p.dim() == 1:
                p.copy_(0.5 + torch.rand(p.shape, generator=g) if "norm" in name and name.endswith("weight")
                        else 0.1 * torch.randn(p.shape, generator=g))
            else:
                p.copy_(torch.randn(p.shape, generator=g) / math.sqrt(p[0].numel()))


@pytest.mark.parametrize("freq,chin,chout,shape", [(True, 4, 48, (2, 4, 64, 21)), (True, 48, 96, (1, 48, 32, 13)),
                                                   (False, 2, 48, (2, 2, 4001)), (False, 96, 192, (1, 96, 515))])
def test_henclayer_is_bit_equal_to_torchaudio(freq, chin, chout, shape):
    a = shim.HEncLayer(chin, chout, freq=freq)
    b = ta._HEncLayer(chin, chout, kernel_size=8, stride=4, freq=freq, norm_type="identity",
                      dconv_kw={"compress": 8, "depth": 2, "init": 1e-3, "norm_type": "group_norm"})
    _seed_module(a, 1)
    missing = b.load_state_dict(a.state_dict(), strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys       # same parameter names and shapes
    assert isinstance(b.norm1, nn.Identity) and isinstance(b.norm2, nn.Identity)
    x = torch.randn(shape, generator=torch.Generator().manual_seed(2))
    with torch.no_grad():
        assert torch.equal(a(x), b(x))


def test_dconv_scaled_embedding_and_rescale_match_torchaudio():
    a, b = shim.DConv(96), ta._DConv(96, compress=8, depth=2, init=1e-3)
    _seed_module(a, 3)
    b.load_state_dict(a.state_dict(), strict=True)
    x = torch.randn(3, 96, 259, generator=torch.Generator().manual_seed(4))
    with torch.no_grad():
        assert torch.equal(a(x), b(x))
    torch.manual_seed(5)
    ea = shim.ScaledEmbedding(512, 48, scale=10.0, smooth=True)
    torch.manual_seed(5)
    eb = ta._ScaledEmbedding(512, 48, scale=10.0, smooth=True)
    assert torch.equal(ea.embedding.weight, eb.embedding.weight)
    idx = torch.arange(512)
    assert torch.equal(ea(idx), eb(idx))
    ma = nn.Sequential(nn.Conv1d(3, 8, 3), nn.ConvTranspose2d(8, 4, (8, 1)), nn.Linear(4, 4))
    mb = nn.Sequential(nn.Conv1d(3, 8, 3), nn.ConvTranspose2d(8, 4, (8, 1)), nn.Linear(4, 4))
    mb.load_state_dict(ma.state_dict())
    shim.rescale_module(ma, 0.1)
    ta._rescale_module(mb)
    for p, q in zip(ma.parameters(), mb.parameters()):
        assert torch.equal(p, q)


@pytest.mark.parametrize("L", [264600, 40000, 5000, 4096])
def test_spectral_front_and_back_end_match_torchaudio(L):
    x = torch.randn(2, 2, L, generator=torch.Generator().manual_seed(L))
    ns = SimpleNamespace(hop_length=1024, nfft=4096)
    ns._pad1d = lambda *a, **k: ta.HDemucs._pad1d(ns, *a, **k)
    z_ta = ta.HDemucs._spec(ns, x)
    z_shim = shim.HTDemucs._spec(ns, x)
    z_or = athtd_oracle.spec(x)
    assert z_ta.shape == (2, 2, 2048, math.ceil(L / 1024))
    assert torch.equal(z_ta, z_shim)      # (inputs shorter than the reflect pad, <= 2559 samples, never reach this path: L >= 4096)
    assert torch.equal(z_shim, z_or)
    assert torch.equal(ta.HDemucs._magnitude(ns, z_ta), athtd_oracle.magnitude(z_ta))
    y_ta = ta.HDemucs._ispec(ns, z_ta, L)
    assert torch.equal(y_ta, shim.HTDemucs._ispec(ns, z_ta, L))
    assert torch.equal(y_ta, athtd_oracle.ispec(z_ta, L))
    assert torch.equal(ta._spectro(x, 4096, 1024), shim.spectro(x, 4096, 1024))


def test_parameter_counts_match_the_reference_dumps():
    """HTDemucs_Fwd_Pass.txt:147 (Total params 41,984,456), :69 (MyTransformerEncoderLayer 3,154,432), :73
    (CrossTransformerEncoderLayer 3,155,456); trainable part / hot-path totals from SURVEY.md section 6."""
    count = lambda m: sum(p.numel() for p in m.parameters())
    assert count(shim.HTDemucs(include_dead_decoders=True)) == 41_984_456
    assert count(shim.MyTransformerEncoderLayer()) == 3_154_432
    assert count(shim.CrossTransformerEncoderLayer()) == 3_155_456
    from oracle import weights
    live = weights.make_state_dict(0)
    assert len(live) == 407
    assert sum(v.numel() for v in live.values()) == 38_196_180            # SURVEY.md section 6 (hot-path parameters)
    h = shim.HTDemucs(include_dead_decoders=False)
    names = {"htdemucs." + k for k, _ in h.named_parameters()}
    assert {k for k in live if k.startswith("htdemucs.")} == names


def test_self_attention_layer_vs_torch_transformer_encoder_layer():
    layer = shim.MyTransformerEncoderLayer().eval()
    _seed_module(layer, 7)
    ref = nn.TransformerEncoderLayer(512, 8, 2048, dropout=0.0, activation=F.gelu, batch_first=True, norm_first=True).eval()
    sd = layer.state_dict()
    g1, g2 = sd["gamma_1.scale"], sd["gamma_2.scale"]
    with torch.no_grad():
        ref.self_attn.in_proj_weight.copy_(sd["self_attn.in_proj_weight"])
        ref.self_attn.in_proj_bias.copy_(sd["self_attn.in_proj_bias"])
        ref.self_attn.out_proj.weight.copy_(g1[:, None] * sd["self_attn.out_proj.weight"])     # LayerScale folded in
        ref.self_attn.out_proj.bias.copy_(g1 * sd["self_attn.out_proj.bias"])
        ref.linear1.weight.copy_(sd["linear1.weight"]); ref.linear1.bias.copy_(sd["linear1.bias"])
        ref.linear2.weight.copy_(g2[:, None] * sd["linear2.weight"]); ref.linear2.bias.copy_(g2 * sd["linear2.bias"])
        ref.norm1.weight.copy_(sd["norm1.weight"]); ref.norm1.bias.copy_(sd["norm1.bias"])
        ref.norm2.weight.copy_(sd["norm2.weight"]); ref.norm2.bias.copy_(sd["norm2.bias"])
    x = torch.randn(2, 77, 512, generator=torch.Generator().manual_seed(8))
    with torch.no_grad():
        y = ref(x)
        # MyGroupNorm(1, 512): statistics over ALL tokens x channels of a sample, per-channel affine
        mu = y.mean(dim=(1, 2), keepdim=True)
        var = y.var(dim=(1, 2), unbiased=False, keepdim=True)
        want = (y - mu) / torch.sqrt(var + 1e-5) * sd["norm_out.weight"] + sd["norm_out.bias"]
        got = layer(x)
        fn = athtd_oracle._self_layer(x, {f"p.{k}": v for k, v in sd.items()}, "p")
    assert (got - want).abs().max() < 2e-5
    assert (fn - got).abs().max() < 2e-5


def test_cross_attention_layer_vs_sdpa_composition():
    layer = shim.CrossTransformerEncoderLayer().eval()
    _seed_module(layer, 9)
    sd = layer.state_dict()
    g = torch.Generator().manual_seed(10)
    q, k = torch.randn(2, 50, 512, generator=g), torch.randn(2, 31, 512, generator=g)
    ln = lambda x, n: F.layer_norm(x, (512,), sd[f"{n}.weight"], sd[f"{n}.bias"], 1e-5)
    W, b = sd["cross_attn.in_proj_weight"], sd["cross_attn.in_proj_bias"]
    with torch.no_grad():
        kn = ln(k, "norm2")
        qq = F.linear(ln(q, "norm1"), W[:512], b[:512]).view(2, 50, 8, 64).transpose(1, 2)
        kk = F.linear(kn, W[512:1024], b[512:1024]).view(2, 31, 8, 64).transpose(1, 2)
        vv = F.linear(kn, W[1024:], b[1024:]).view(2, 31, 8, 64).transpose(1, 2)
        o = F.scaled_dot_product_attention(qq, kk, vv).transpose(1, 2).reshape(2, 50, 512)
        x = q + sd["gamma_1.scale"] * F.linear(o, sd["cross_attn.out_proj.weight"], sd["cross_attn.out_proj.bias"])
        h = F.linear(F.gelu(F.linear(ln(x, "norm3"), sd["linear1.weight"], sd["linear1.bias"])), sd["linear2.weight"], sd["linear2.bias"])
        x = x + sd["gamma_2.scale"] * h
        want = F.group_norm(x.transpose(1, 2), 1, sd["norm_out.weight"], sd["norm_out.bias"], 1e-5).transpose(1, 2)
        got = layer(q, k)
        fn = athtd_oracle._cross_layer(q, k, {f"p.{k_}": v for k_, v in sd.items()}, "p")
    assert (got - want).abs().max() < 2e-5
    assert (fn - got).abs().max() < 2e-5


def test_sinusoidal_embeddings_closed_form():
    """demucs transformer.py create_sin_embedding / create_2d_sin_embedding in float64."""
    T, C = 1034, 512
    pe = shim.create_sin_embedding(T, C)[:, 0].double().numpy()
    pos = np.arange(T)[:, None]
    ph = pos / (10000.0 ** (np.arange(C // 2) / (C // 2 - 1)))
    assert np.abs(pe - np.concatenate([np.cos(ph), np.sin(ph)], axis=1)).max() < 2e-4      # fp32 phase rounding at t ~ 1000
    H, Wd = 8, 259
    pe2 = shim.create_2d_sin_embedding(C, H, Wd)[0].double().numpy()                        # [C, H, W]
    d = C // 2
    div = np.exp(np.arange(0, d, 2) * -(math.log(10000.0) / d))
    w, h = np.arange(Wd)[:, None] * div, np.arange(H)[:, None] * div
    want = np.zeros((C, H, Wd))
    want[0:d:2] = np.sin(w).T[:, None, :]
    want[1:d:2] = np.cos(w).T[:, None, :]
    want[d::2] = np.sin(h).T[:, :, None]
    want[d + 1::2] = np.cos(h).T[:, :, None]
    assert np.abs(pe2 - want).max() < 1e-4


def test_functional_cross_transformer_matches_the_shim_module():
    xf = shim.CrossTransformerEncoder().eval()
    _seed_module(xf, 11)
    sd = {"htdemucs.crosstransformer." + k: v for k, v in xf.state_dict().items()}
    g = torch.Generator().manual_seed(12)
    x, xt = torch.randn(1, 512, 8, 9, generator=g), torch.randn(1, 512, 37, generator=g)
    with torch.no_grad():
        a, at = xf(x, xt)
        b, bt = athtd_oracle.cross_transformer(sd, x, xt)
    assert (a - b).abs().max() < 5e-5 and (at - bt).abs().max() < 5e-5
