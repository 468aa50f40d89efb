"""TEST INFRASTRUCTURE ONLY -- CPU restatement (torch, fp32) of the CLAP text tower this path calls
(/root/reference/src/models/stem_separation/ATHTDemucs_v2.py:238-248 -> HF transformers, pinned 4.51.1 in the reference's
requirements.txt:14; structure printed at src/models/stem_separation/AudioTextHTDemucs_Full.txt:630-823):
ClapTextModelWithProjection = RoBERTa-base embeddings + 12 post-LN encoder layers + tanh pooler + ClapProjectionLayer.
Pinned against the transformers implementation installed in the build container by oracle/make_clap_fixtures.py
(fixtures: tests/golden/clap_text.json)."""
import math

import torch
import torch.nn.functional as F

H, LAYERS, HEADS, FF, VOCAB, NPOS, PAD, PROJ = 768, 12, 12, 3072, 50265, 514, 1, 512


def param_shapes():
    s = {"text_model.embeddings.word_embeddings.weight": (VOCAB, H), "text_model.embeddings.position_embeddings.weight": (NPOS, H),
         "text_model.embeddings.token_type_embeddings.weight": (1, H), "text_model.embeddings.LayerNorm.weight": (H,),
         "text_model.embeddings.LayerNorm.bias": (H,)}
    for l in range(LAYERS):
        p = f"text_model.encoder.layer.{l}."
        for n in ("attention.self.query", "attention.self.key", "attention.self.value", "attention.output.dense"):
            s[p + n + ".weight"] = (H, H); s[p + n + ".bias"] = (H,)
        s[p + "attention.output.LayerNorm.weight"] = (H,); s[p + "attention.output.LayerNorm.bias"] = (H,)
        s[p + "intermediate.dense.weight"] = (FF, H); s[p + "intermediate.dense.bias"] = (FF,)
        s[p + "output.dense.weight"] = (H, FF); s[p + "output.dense.bias"] = (H,)
        s[p + "output.LayerNorm.weight"] = (H,); s[p + "output.LayerNorm.bias"] = (H,)
    s["text_model.pooler.dense.weight"] = (H, H); s["text_model.pooler.dense.bias"] = (H,)
    s["text_projection.linear1.weight"] = (PROJ, H); s["text_projection.linear1.bias"] = (PROJ,)
    s["text_projection.linear2.weight"] = (PROJ, PROJ); s["text_projection.linear2.bias"] = (PROJ,)
    return s


def make_state_dict(seed: int = 0):
    """Seeded weights with trained-model-like scales (N(0, 0.05) matrices, LayerNorm weights around 1, small biases) so that
    every branch matters numerically."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, shp in param_shapes().items():
        if k.endswith("LayerNorm.weight"):
            sd[k] = 1.0 + 0.1 * torch.randn(shp, generator=g)
        elif k.endswith(".bias"):
            sd[k] = 0.05 * torch.randn(shp, generator=g)
        else:
            sd[k] = 0.05 * torch.randn(shp, generator=g)
    return sd


def make_inputs(seed: int, P: int, S: int):
    """Random token ids with RoBERTa framing (<s>=0 ... </s>=2, pad=1) and ragged lengths."""
    g = torch.Generator().manual_seed(seed)
    ids = torch.full((P, S), PAD, dtype=torch.long)
    mask = torch.zeros(P, S, dtype=torch.long)
    for p in range(P):
        n = S if p == 0 else int(torch.randint(3, S + 1, (1,), generator=g))
        ids[p, 0] = 0
        ids[p, 1:n - 1] = torch.randint(3, VOCAB, (n - 2,), generator=g)
        ids[p, n - 1] = 2
        mask[p, :n] = 1
    return ids, mask


def forward(sd, input_ids: torch.Tensor, attention_mask: torch.Tensor, normalize: bool = False) -> torch.Tensor:
    """ClapTextModelWithProjection.forward(...).text_embeds (normalize=False) / ClapModel.get_text_features (normalize=True):
    modeling_clap.py ClapTextEmbeddings (position ids = cumsum(ids != pad) * (ids != pad) + pad), ClapTextLayer x 12
    (post-LayerNorm, eps 1e-12, exact GELU), ClapTextPooler (tanh of dense on token 0), ClapProjectionLayer (linear, ReLU, linear)."""
    e = "text_model.embeddings."
    nonpad = input_ids.ne(PAD).int()
    pos = (torch.cumsum(nonpad, dim=1) * nonpad).long() + PAD
    x = sd[e + "word_embeddings.weight"][input_ids] + sd[e + "token_type_embeddings.weight"][0] + sd[e + "position_embeddings.weight"][pos]
    x = F.layer_norm(x, (H,), sd[e + "LayerNorm.weight"], sd[e + "LayerNorm.bias"], 1e-12)
    P, S = input_ids.shape
    add = (1.0 - attention_mask[:, None, None, :].float()) * torch.finfo(torch.float32).min
    for l in range(LAYERS):
        p = f"text_model.encoder.layer.{l}."
        lin = lambda t, n: F.linear(t, sd[p + n + ".weight"], sd[p + n + ".bias"])
        q = lin(x, "attention.self.query").view(P, S, HEADS, 64).transpose(1, 2)
        k = lin(x, "attention.self.key").view(P, S, HEADS, 64).transpose(1, 2)
        v = lin(x, "attention.self.value").view(P, S, HEADS, 64).transpose(1, 2)
        pr = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(64) + add, dim=-1)
        ctx = (pr @ v).transpose(1, 2).reshape(P, S, H)
        x = F.layer_norm(lin(ctx, "attention.output.dense") + x, (H,), sd[p + "attention.output.LayerNorm.weight"],
                         sd[p + "attention.output.LayerNorm.bias"], 1e-12)
        h = F.gelu(lin(x, "intermediate.dense"))
        x = F.layer_norm(lin(h, "output.dense") + x, (H,), sd[p + "output.LayerNorm.weight"], sd[p + "output.LayerNorm.bias"], 1e-12)
    pooled = torch.tanh(F.linear(x[:, 0], sd["text_model.pooler.dense.weight"], sd["text_model.pooler.dense.bias"]))
    out = F.linear(F.relu(F.linear(pooled, sd["text_projection.linear1.weight"], sd["text_projection.linear1.bias"])),
                   sd["text_projection.linear2.weight"], sd["text_projection.linear2.bias"])
    return F.normalize(out, dim=-1) if normalize else out
