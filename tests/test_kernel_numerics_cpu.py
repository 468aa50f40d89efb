"""CPU restatement of the arithmetic claims behind the attention softmax (csrc/attention.cu), with the constants read from the source:
the FMA-pipe exp2 (Cody-Waite + degree-3 polynomial) and the optimistic fixed-reference pass with its validity window."""
import os
import re

import numpy as np

SRC = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "audio-to-sheet-music_b200", "csrc", "attention.cu")).read()


def _const(name):
    m = re.search(name + r"\s*=\s*([0-9.eE+-]+)f", SRC)
    assert m, name
    return np.float32(m.group(1))


def _poly_coeffs():
    # q = fma(f, c3, c2); q = fma(q, f, c1); q = fma(q, f, c0)
    m = re.search(r"f2fma\(f, make_float2\(([0-9.]+)f, [0-9.]+f\), make_float2\(([0-9.]+)f", SRC)
    c3, c2 = np.float32(m.group(1)), np.float32(m.group(2))
    rest = re.findall(r"q = f2fma\(q, f, make_float2\(([0-9.]+)f", SRC)
    return c3, c2, np.float32(rest[0]), np.float32(rest[1])


def _exp2_poly(a):
    """The kernel's sequence in float32: clamp, round through the 1.5 * 2^23 magic constant, polynomial, exponent add."""
    c3, c2, c1, c0 = _poly_coeffs()
    a = np.minimum(np.maximum(a.astype(np.float32), np.float32(-125.0)), np.float32(127.0))
    magic = np.float32(12582912.0)
    t = (a + magic).astype(np.float32)
    f = (a - (t - magic).astype(np.float32)).astype(np.float32)
    q = (f * c3 + c2).astype(np.float32)
    q = (q * f + c1).astype(np.float32)
    q = (q * f + c0).astype(np.float32)
    bits = (q.view(np.uint32) + (t.view(np.uint32) << np.uint32(23))).astype(np.uint32)
    return bits.view(np.float32)


def test_polynomial_exp2_matches_exp2_to_the_stated_accuracy():
    rng = np.random.default_rng(0)
    a = np.concatenate([rng.uniform(-125, 127, 200000), np.arange(-125, 128, dtype=np.float64), rng.uniform(-2, 2, 50000)]).astype(np.float32)
    got = _exp2_poly(a).astype(np.float64)
    ref = np.exp2(a.astype(np.float64))
    rel = np.abs(got - ref) / ref
    assert rel.max() < 8e-5, rel.max()          # attention.cu: "relative error 7.5e-5"
    # out-of-range arguments saturate instead of wrapping the exponent field
    assert _exp2_poly(np.array([-1e4, -np.inf], np.float32)).max() <= 2.0 ** -124
    assert _exp2_poly(np.array([500.0, 1e6], np.float32)).min() >= 2.0 ** 126


def _bf16(x):
    u = x.astype(np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32)


def _optimistic_softmax_pv(s, v, cs):
    """One row block: reference = max of the first 64 keys + headroom; P in fp32 (ftz), bf16 for the PV product, fp32 row sum."""
    head, lmax = _const("FA_HEADROOM"), _const("FA_LSUM_MAX")
    mb = (s[:, :64].max(axis=1, keepdims=True) * cs + head).astype(np.float32)
    a = (s * cs - mb).astype(np.float32)
    with np.errstate(over="ignore"):
        p = np.exp2(a.astype(np.float64)).astype(np.float32)
    p[p < np.float32(2.0 ** -126)] = 0.0            # ex2.approx.ftz
    l = p.sum(axis=1, dtype=np.float32)
    valid = l < lmax
    with np.errstate(over="ignore", invalid="ignore"):
        o = (_bf16(p).astype(np.float64) @ v.astype(np.float64)) / l[:, None].astype(np.float64)
    return o, valid


def test_optimistic_reference_is_exact_inside_its_window_and_flags_outside():
    rng = np.random.default_rng(1)
    cs = np.float32(0.125 * 1.4426950408889634)
    rows, keys = 64, 1034
    v = rng.standard_normal((keys, 64)).astype(np.float32)
    for jump_nats in (0.0, 40.0, 100.0):             # logits of the late keys exceed the first tile's maximum by this much
        s = (8.0 * rng.standard_normal((rows, keys))).astype(np.float32)        # raw scores (logit = s / 8)
        s[:, 700] += np.float32(8.0 * jump_nats)
        o, valid = _optimistic_softmax_pv(s, v, cs)
        assert valid.all(), jump_nats
        logits = s.astype(np.float64) / 8.0
        w = np.exp(logits - logits.max(axis=1, keepdims=True))
        ref = (w / w.sum(axis=1, keepdims=True)) @ v.astype(np.float64)
        assert np.abs(o - ref).max() < 1e-2 + 1e-2 * np.abs(ref).max(), jump_nats
    # beyond the window (160 binades = 111 nats above the first tile's maximum) the row sum must leave [0, 2^100): the CTA repeats
    s = (8.0 * rng.standard_normal((rows, keys))).astype(np.float32)
    s[:, 700] += np.float32(8.0 * 130.0)
    _, valid = _optimistic_softmax_pv(s, v, cs)
    assert not valid.any()
    head, lmax = float(_const("FA_HEADROOM")), float(_const("FA_LSUM_MAX"))
    assert head == 60.0 and lmax == 2.0 ** 100


def test_fast_gelu_fit_matches_erf_gelu():
    """common.cuh gelu_fast: erf(x / sqrt 2) ~ tanh(x (a0 + a1 x^2 + a2 x^4)) with x^2 clamped at 32; the header states |GELU error| <= 2.9e-5
    with an exact tanh (tanh.approx adds a relative 2^-11 on top, covered by the GPU parity tests)."""
    from math import erf
    src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "audio-to-sheet-music_b200", "csrc", "common.cuh")).read()
    body = src[src.index("float gelu_fast(float x)"):]
    a2, a1 = [float(v) for v in re.search(r"fmaf\((-?[0-9.e+-]+)f, x2, ([0-9.e+-]+)f\)", body).groups()]
    a0 = float(re.search(r"fmaf\(p, x2, ([0-9.e+-]+)f\)", body).group(1))
    x = np.linspace(-12.0, 12.0, 480001)
    x2 = np.minimum(x * x, 32.0)
    fast = 0.5 * x * (1.0 + np.tanh(x * ((a2 * x2 + a1) * x2 + a0)))
    ref = 0.5 * x * (1.0 + np.vectorize(erf)(x / np.sqrt(2.0)))
    assert np.abs(fast - ref).max() <= 2.9e-5 * 1.05
