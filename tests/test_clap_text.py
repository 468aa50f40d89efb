"""CLAP text tower.  CPU: the oracle restatement against fixtures produced by the real HF
ClapTextModelWithProjection (oracle/make_clap_fixtures.py).  GPU: the native tower (C ABI) against the oracle and the
fixtures (fp32, tolerance 2e-4 max-abs on O(1) embeddings), and end to end as the ``clap_encoder`` of the module mirror."""
import json
import os

import pytest
import torch

from oracle import clap_text as oc

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = json.load(open(os.path.join(HERE, "golden", "clap_text.json")))


@pytest.fixture(scope="module")
def clap_sd():
    return oc.make_state_dict(0)


@pytest.mark.parametrize("i", range(len(CASES)))
def test_oracle_tower_matches_transformers_fixture(clap_sd, i):
    c = CASES[i]
    ids, mask = oc.make_inputs(c["seed"], c["P"], c["S"])
    out = oc.forward(clap_sd, ids, mask)
    assert (out - torch.tensor(c["text_embeds"])).abs().max() < 1e-4


def test_native_param_table_covers_the_hf_state_dict():
    from athtd_b200 import lib as alib
    lib = alib.load()
    names = {lib.athtd_clap_param_name(i).decode(): lib.athtd_clap_param_numel(i) for i in range(lib.athtd_clap_param_count())}
    shapes = oc.param_shapes()
    assert set(names) == set(shapes)
    for k, shp in shapes.items():
        n = 1
        for d in shp:
            n *= d
        assert names[k] == n


@pytest.mark.gpu
@pytest.mark.parametrize("i", range(len(CASES)))
def test_native_tower_matches_oracle_and_fixture(clap_sd, i):
    import athtd_b200
    c = CASES[i]
    ids, mask = oc.make_inputs(c["seed"], c["P"], c["S"])
    tower = athtd_b200.ClapModelTextB200()
    tower.load_state_dict(clap_sd)
    tower = tower.cuda()
    out = tower(input_ids=ids.cuda(), attention_mask=mask.cuda()).text_embeds.cpu()
    assert (out - torch.tensor(c["text_embeds"])).abs().max() < 2e-4
    feat = tower.get_text_features(input_ids=ids.cuda(), attention_mask=mask.cuda()).cpu()
    assert (feat - oc.forward(clap_sd, ids, mask, normalize=True)).abs().max() < 2e-5
    assert torch.allclose(feat.norm(dim=-1), torch.ones(c["P"]), atol=1e-5)


@pytest.mark.gpu
def test_module_forward_with_native_clap(clap_sd, state_dict):
    """forward(wav, List[str]) with the native tower as clap_encoder and a stand-in tokenizer (the real one is a vocabulary
    file): same result as passing the oracle tower's embeddings as a tensor; CLAP runs once per distinct prompt (cache)."""
    import athtd_b200
    from oracle import athtd_oracle, weights

    class Tok:
        calls = 0

        def __call__(self, text, padding=True, return_tensors="pt"):
            Tok.calls += 1
            ids, mask = oc.make_inputs(100 + len(text), len(text), 7)
            return {"input_ids": ids, "attention_mask": mask}

    tower = athtd_b200.ClapModelTextB200()
    tower.load_state_dict(clap_sd)
    m = athtd_b200.AudioTextHTDemucsB200(None, tower, Tok(), precision="fp32")
    m.load_state_dict(state_dict, strict=False)
    m = m.cuda().eval()
    wav, _ = weights.make_inputs(5, 2, 20000)
    out = m(wav.cuda(), ["drums", "bass"]).cpu()
    again = m(wav.cuda(), ["drums", "bass"]).cpu()
    assert Tok.calls == 1 and torch.equal(out, again)
    ids, mask = oc.make_inputs(102, 2, 7)
    emb = oc.forward(clap_sd, ids, mask, normalize=True)
    ref = athtd_oracle.forward(state_dict, wav, emb)
    assert (out - ref).abs().max() < 1e-3
