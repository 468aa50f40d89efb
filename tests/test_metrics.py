"""Evaluation metrics.  CPU: the oracle restatement against fixtures produced by the reference's own
src/loss.py (oracle/make_metric_fixtures.py).  GPU: athtd_b200.metrics (one reduction kernel through the C ABI) against
the oracle; tolerance 2e-3 dB (fp64 moment expansion vs the reference's fp32 element-wise evaluation)."""
import json
import os

import pytest
import torch

from oracle import metrics as om

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = json.load(open(os.path.join(HERE, "golden", "metrics.json")))


@pytest.mark.parametrize("i", range(len(CASES)))
def test_oracle_metrics_match_reference_fixtures(i):
    c = CASES[i]
    est, tgt = om.make_case(**c["case"])
    assert abs(float(om.sdr_loss(est, tgt)) - c["sdr_loss"]) < 1e-4
    assert abs(float(om.sisdr_loss(est, tgt)) - c["sisdr_loss"]) < 1e-4
    got = [float(v) for v in om.new_sdr_metric(est, tgt)]
    assert max(abs(a - b) for a, b in zip(got, c["new_sdr_metric"])) < 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize("i", range(len(CASES)))
def test_device_metrics_match_oracle_and_fixtures(i):
    import athtd_b200
    from athtd_b200 import metrics as dm
    c = CASES[i]
    est, tgt = om.make_case(**c["case"])
    e, t = est.cuda(), tgt.cuda()
    assert abs(float(dm.sdr_loss(e, t)) - c["sdr_loss"]) < 2e-3
    assert abs(float(dm.sisdr_loss(e, t)) - c["sisdr_loss"]) < 2e-3
    got = dm.new_sdr_metric(e, t).cpu()
    ref = torch.tensor(c["new_sdr_metric"])
    big = ref > 100                                   # noise-free case: 10 log10(sum t^2 / 1e-8), fp32-sum sensitive
    assert (got[~big] - ref[~big]).abs().max() < 2e-3 if (~big).any() else True
    assert ((got[big] - ref[big]).abs() < 0.05).all()
    assert abs(dm.compute_sdr(e[0], t[0]) - float(-om.sdr_loss(est[:1], tgt[:1]))) < 2e-3
    assert abs(dm.compute_sisdr(e[0], t[0]) - float(-om.sisdr_loss(est[:1], tgt[:1]))) < 2e-3
    with pytest.raises(athtd_b200.AthtdError):
        dm.sdr_loss(est, tgt)                          # CPU tensors: no fallback
