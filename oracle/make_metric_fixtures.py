"""Run IN THE BUILD CONTAINER ONLY (needs /root/reference): evaluates the reference's own src/loss.py functions on the
seeded cases of oracle/metrics.make_case and writes tests/golden/metrics.json; also asserts the oracle restatement agrees."""
import importlib.util
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import metrics as om

spec = importlib.util.spec_from_file_location("ref_loss", "/root/reference/src/loss.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

CASES = [dict(seed=1, batch=3, channels=2, n=44100, noise=0.05), dict(seed=2, batch=2, channels=2, n=264600, noise=0.3, gain=0.5),
         dict(seed=3, batch=1, channels=2, n=10007, noise=0.01, gain=1.0, offset=0.2), dict(seed=4, batch=4, channels=1, n=5000, noise=2.0),
         dict(seed=5, batch=1, channels=2, n=30000, noise=0.0)]
out = []
for c in CASES:
    est, tgt = om.make_case(**c)
    r = {"case": c, "sdr_loss": float(ref.sdr_loss(est, tgt)), "sisdr_loss": float(ref.sisdr_loss(est, tgt)),
         "new_sdr_metric": [float(v) for v in ref.new_sdr_metric(est, tgt)]}
    assert abs(float(om.sdr_loss(est, tgt)) - r["sdr_loss"]) < 1e-5
    assert abs(float(om.sisdr_loss(est, tgt)) - r["sisdr_loss"]) < 1e-5
    assert max(abs(a - b) for a, b in zip([float(v) for v in om.new_sdr_metric(est, tgt)], r["new_sdr_metric"])) < 1e-5
    out.append(r)
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "metrics.json"), "w"), indent=1)
print("wrote", len(out), "cases")
