"""Run IN THE BUILD CONTAINER (needs the `transformers` package): loads the seeded weights of oracle/clap_text.py into the real
HF ClapTextModelWithProjection (the class the reference instantiates, ATHTDemucs_v2.py:19), runs it on the seeded token batches
and writes tests/golden/clap_text.json; asserts that the oracle restatement agrees with it."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from transformers import ClapTextConfig
from transformers.models.clap.modeling_clap import ClapTextModelWithProjection
from oracle import clap_text as oc

sd = oc.make_state_dict(0)
hf = ClapTextModelWithProjection(ClapTextConfig()).eval()
res = hf.load_state_dict(sd, strict=False)
assert not res.unexpected_keys and all("position_ids" in k or "token_type_ids" in k for k in res.missing_keys), res
out = []
for (seed, P, S) in [(1, 3, 9), (2, 4, 6), (3, 1, 16)]:
    ids, mask = oc.make_inputs(seed, P, S)
    with torch.no_grad():
        ref = hf(input_ids=ids, attention_mask=mask).text_embeds
    mine = oc.forward(sd, ids, mask)
    err = float((ref - mine).abs().max())
    assert err < 2e-5, err
    out.append({"seed": seed, "P": P, "S": S, "text_embeds": ref.tolist(), "oracle_vs_hf_max_abs": err})
    print(f"case seed={seed} P={P} S={S}: oracle vs transformers max-abs {err:.2e}")
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "clap_text.json"), "w"))
