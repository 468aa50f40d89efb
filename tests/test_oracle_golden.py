"""CPU: the oracle restatement against the fixtures generated in the build container from the REAL
reference module (oracle/validate_against_reference.py).  Tolerance: fp32, max-abs 1e-4."""
import json
import os

import numpy as np
import torch

from oracle import athtd_oracle, ola, weights


def test_state_dict_checksum(state_dict, golden_dir):
    g = np.load(os.path.join(golden_dir, "forward_short.npz"))
    assert abs(weights.state_dict_checksum(state_dict) - float(g["sd_checksum"])) < 1e-3 * float(g["sd_checksum"]) * 1e-6 + 1e-2


def test_forward_short_matches_reference_fixture(state_dict, golden_dir):
    g = np.load(os.path.join(golden_dir, "forward_short.npz"))
    wav, emb = weights.make_inputs(11, 2, 40000)
    out = athtd_oracle.forward(state_dict, wav, emb)
    assert out.shape == (2, 2, 40000)
    assert np.abs(out.numpy() - g["out"]).max() < 1e-4
    wav, emb = weights.make_inputs(12, 1, 30000, emb_norm=False)      # un-normalised embedding (quirk Q7)
    out = athtd_oracle.forward(state_dict, wav, emb)
    assert np.abs(out.numpy() - g["out_unnorm"]).max() < 1e-4


def test_forward_6s_matches_reference_fixture(state_dict, golden_dir):
    g = np.load(os.path.join(golden_dir, "forward_6s.npz"))
    wav, emb = weights.make_inputs(1, 1, 264600)
    taps = {}
    out = athtd_oracle.forward(state_dict, wav, emb, taps)
    assert np.abs(out[..., ::37].numpy() - g["out_dec37"]).max() < 1e-4
    assert abs(float((out.double() ** 2).sum()) - float(g["out_sumsq"])) < 1e-3 * float(g["out_sumsq"])
    assert np.abs(taps["x_dec"].numpy() - g["x_dec"]).max() < 1e-4
    # shapes of SURVEY.md Appendix B
    assert taps["enc3"].shape == (1, 384, 8, 259) and taps["tenc3"].shape == (1, 384, 1034)
    assert taps["xf_layer4"].shape == (1, 2072, 512) and taps["xf_layer4_t"].shape == (1, 1034, 512)


def test_dc_imag_is_exactly_zero():
    wav, _ = weights.make_inputs(3, 1, 20000)
    z = athtd_oracle.spec(wav)
    assert (z[:, :, 0, :].imag == 0).all()


def test_chunk_plans_match_reference_fixture(golden_dir):
    g = json.load(open(os.path.join(golden_dir, "ola_plans.json")))
    for T, rec in g.items():
        plan = [list(map(int, c[:4])) + [int(c.fade_in), int(c.fade_out)] for c in ola.chunk_plan(int(T))]
        assert plan == rec["plan"], T
        gen = torch.Generator().manual_seed(int(T))
        mix = torch.randn(2, int(T), generator=gen)
        out = ola.chunked_inference(lambda c: torch.tanh(c * 3.0) + 0.25 * c, mix)
        assert abs(float(out.double().sum()) - rec["sum"]) <= 1e-9 * max(1.0, abs(rec["sum"]))
        assert abs(float((out.double() ** 2).sum()) - rec["sumsq"]) <= 1e-9 * max(1.0, rec["sumsq"])


def test_config3_edge_case():
    """4-minute track -> 54 chunks; #52 ends exactly at T (no fade-out), #53 is a 66150-sample tail."""
    plan = ola.chunk_plan(240 * 44100)
    assert len(plan) == 54
    assert plan[52].end == 240 * 44100 and not plan[52].fade_out and plan[52].fade_in
    assert plan[53].actual_len == 66150 and plan[53].fade_len == 33075 and plan[53].fade_in and not plan[53].fade_out
