"""TEST INFRASTRUCTURE -- pins the oracle against the real reference, in THIS container only.

Run:  python -m oracle.validate_against_reference [--write-golden]

What it does
  1. Imports /root/reference/src/models/stem_separation/ATHTDemucs_v2.py UNMODIFIED, with
     ``demucs.htdemucs.HTDemucs`` stubbed by oracle/demucs_shim.py (demucs 4.0.1 is not
     installable here) and ``_get_clap_embeddings`` (ATHTDemucs_v2.py:238-248) replaced by a
     function returning the synthetic (B,512) embedding.
  2. Loads the seeded state_dict of oracle/weights.py into it (checks the key layout) and
     compares its forward with the functional restatement oracle/athtd_oracle.py.
  3. Extracts ``OurModel._chunked_inference`` from /root/reference/benchmark.py:155-204 with
     ``ast`` (benchmark.py itself cannot be imported: demucs/matplotlib missing), executes that
     function verbatim with a stand-in model, and compares oracle/ola.py with it bit-for-bit.
  4. With --write-golden, writes the fixtures under tests/golden/ that the CPU and GPU tests
     check on machines where /root/reference does not exist.
Nothing from the reference is copied into the repository: only numeric outputs are stored.
"""
from __future__ import annotations

import argparse
import ast
import importlib.util
import json
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import athtd_oracle, demucs_shim, ola, weights

REF = "/root/reference"
GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def import_reference_model():
    demucs = types.ModuleType("demucs")
    demucs_ht = types.ModuleType("demucs.htdemucs")
    demucs_ht.HTDemucs = demucs_shim.HTDemucs
    demucs.htdemucs = demucs_ht
    sys.modules.setdefault("demucs", demucs)
    sys.modules.setdefault("demucs.htdemucs", demucs_ht)
    path = os.path.join(REF, "src/models/stem_separation/ATHTDemucs_v2.py")
    spec = importlib.util.spec_from_file_location("ref_athtdemucs_v2", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def build_reference(sd):
    mod = import_reference_model()
    ht = demucs_shim.HTDemucs(include_dead_decoders=True)
    model = mod.AudioTextHTDemucs(ht, nn.Module(), None)
    missing, unexpected = model.load_state_dict(sd, strict=False)
    live_missing = [k for k in missing if not (k.startswith("htdemucs.decoder.") or k.startswith("htdemucs.tdecoder."))]
    assert not live_missing, f"oracle key layout misses live reference keys: {live_missing[:5]}"
    assert not unexpected, f"oracle has keys the reference does not: {unexpected[:5]}"
    model.eval()
    return model


def reference_forward(model, wav, emb):
    model._get_clap_embeddings = lambda text, device: emb
    with torch.no_grad():
        return model(wav, ["synthetic"] * wav.shape[0])


def extract_reference_chunk_loop():
    src = open(os.path.join(REF, "benchmark.py")).read()
    tree = ast.parse(src)
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name == "OurModel":
            for fn in node.body:
                if isinstance(fn, ast.FunctionDef) and fn.name == "_chunked_inference":
                    m = ast.Module(body=[fn], type_ignores=[])
                    ns = {"torch": torch, "F": F, "SAMPLE_RATE": 44100}
                    exec(compile(m, "benchmark.py:_chunked_inference", "exec"), ns)
                    return ns["_chunked_inference"]
    raise RuntimeError("reference chunk loop not found")


def standin_model(chunk, prompts=None):
    return torch.tanh(chunk * 3.0) + 0.25 * chunk


def check_forward(sd, seed, batch, length, emb_norm=True):
    wav, emb = weights.make_inputs(seed, batch, length, emb_norm)
    model = build_reference(sd)
    ref = reference_forward(model, wav, emb)
    taps = {}
    mine = athtd_oracle.forward(sd, wav, emb, taps)
    err = float((ref - mine).abs().max())
    print(f"forward B={batch} L={length}: ref-vs-restatement max-abs {err:.3e}, |ref|max {float(ref.abs().max()):.3f}")
    assert err < 2e-4, err
    return wav, emb, ref, taps


def check_ola():
    ref_loop = extract_reference_chunk_loop()
    out = {}
    for T in (1, 66149, 66150, 132301, 264600, 264601, 463050, 600000, 1000003):
        g = torch.Generator().manual_seed(T)
        mix = torch.randn(2, T, generator=g)
        fake_self = types.SimpleNamespace(device="cpu", segment_seconds=6.0, overlap=1.5, model=standin_model)
        ref = ref_loop(fake_self, mix, "drums")
        mine = ola.chunked_inference(lambda c: standin_model(c), mix)
        assert torch.equal(ref, mine), f"OLA restatement differs from reference at T={T}"
        out[str(T)] = {
            "plan": [list(map(int, c[:4])) + [int(c.fade_in), int(c.fade_out)] for c in ola.chunk_plan(T)],
            "sum": float(ref.double().sum()), "sumsq": float((ref.double() ** 2).sum()),
        }
    print("OLA restatement == reference loop (bit-exact) for", list(out))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--write-golden", action="store_true")
    args = ap.parse_args()
    torch.set_num_threads(os.cpu_count())
    sd = weights.make_state_dict(0)
    ola_golden = check_ola()
    wav_s, emb_s, ref_s, taps_s = check_forward(sd, 11, 2, 40000)
    wav_u, emb_u, ref_u, _ = check_forward(sd, 12, 1, 30000, emb_norm=False)
    wav_6, emb_6, ref_6, taps_6 = check_forward(sd, 1, 1, 264600)
    if args.write_golden:
        os.makedirs(GOLDEN, exist_ok=True)
        np.savez(os.path.join(GOLDEN, "forward_short.npz"), out=ref_s.numpy(), out_unnorm=ref_u.numpy(),
                 sd_checksum=weights.state_dict_checksum(sd),
                 wav_checksum=float(wav_s.double().abs().sum()))
        tap_stats = {k: [float(v.abs().double().mean()) if not v.is_complex() else float(v.abs().double().mean()),
                         float(v.abs().max())] for k, v in taps_6.items()}
        np.savez(os.path.join(GOLDEN, "forward_6s.npz"), out_dec37=ref_6[..., ::37].numpy(),
                 out_sum=float(ref_6.double().sum()), out_sumsq=float((ref_6.double() ** 2).sum()),
                 x_dec=taps_6["x_dec"].numpy(), tap_stats=json.dumps(tap_stats),
                 sd_checksum=weights.state_dict_checksum(sd),
                 wav_checksum=float(wav_6.double().abs().sum()))
        with open(os.path.join(GOLDEN, "ola_plans.json"), "w") as f:
            json.dump(ola_golden, f)
        print("golden fixtures written to", GOLDEN)
        for k, v in tap_stats.items():
            print(f"  tap {k:32s} mean|x| {v[0]:.4f} max|x| {v[1]:.3f}")


if __name__ == "__main__":
    main()
