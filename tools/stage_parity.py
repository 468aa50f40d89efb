"""Per-stage parity report: CUDA taps vs the CPU oracle's taps on the same seeded inputs.
Usage: python tools/stage_parity.py [--precision fp32|bf16] [--B 2] [--L 40000] [--P 1]"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import athtd_b200
from oracle import athtd_oracle, weights


def interior(t, dims):
    G, Rp, C, pf = dims
    return t.view(G, Rp, C)


def report(name, ours, ref):
    ours = ours.float().cpu()
    ref = ref.float()
    if ours.shape != ref.shape:
        print(f"{name:34s} SHAPE MISMATCH ours {tuple(ours.shape)} ref {tuple(ref.shape)}")
        return
    err = (ours - ref).abs().max().item()
    scale = ref.abs().max().item()
    snr = athtd_oracle.snr_db(ours, ref)
    bad = "" if torch.isfinite(ours).all() else "  NON-FINITE"
    print(f"{name:34s} max-abs {err:10.3e}  ref-max {scale:8.3f}  SNR {snr:7.2f} dB{bad}")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--precision", default="fp32")
    ap.add_argument("--B", type=int, default=2)
    ap.add_argument("--L", type=int, default=40000)
    ap.add_argument("--P", type=int, default=1)
    a = ap.parse_args()
    torch.set_num_threads(os.cpu_count())
    sd = weights.make_state_dict(0)
    wav, emb = weights.make_inputs(11, a.B, a.L)
    embs = [emb] + [weights.make_inputs(100 + p, a.B, 4096)[1] for p in range(1, a.P)]
    taps = {}
    t0 = time.time()
    ref = athtd_oracle.forward(sd, wav, embs[0], taps)
    print(f"oracle forward {time.time() - t0:.2f}s")
    model = athtd_b200.AudioTextHTDemucsB200(precision=a.precision)
    model.load_state_dict(sd, strict=False)
    model = model.cuda().eval()
    E = torch.stack(embs, dim=1).cuda()
    out = model.separate_batch(wav.cuda(), E)
    torch.cuda.synchronize()
    plan = model.engine().plan(a.B, a.L, a.P)
    print("launches per forward:", plan.launches, "tcgen05 GEMM launches:", plan.tc_launches)
    B, L = a.B, a.L
    Tf = plan.Tf

    def rs(name):
        return plan.tap(name).interior(), 0

    z = plan.tap("Z").to_torch().view(B, Tf, 2048, 4).cpu()
    zr = torch.view_as_real(taps["z"])          # [B,2,F,T,2]
    zref = torch.stack([zr[:, 0, :, :, 0], zr[:, 0, :, :, 1], zr[:, 1, :, :, 0], zr[:, 1, :, :, 1]], dim=-1).permute(0, 2, 1, 3)
    report("Z (stft cac)", z, zref)
    t, pf = rs("xf0"); report("xf0 (norm spec)", t[:, pf:pf + 2048].reshape(B, Tf, 2048, 4), taps["x_norm"].permute(0, 3, 2, 1))
    t, pf = rs("xt0"); report("xt0 (norm wav)", t[:, pf:pf + L], taps["xt_norm"].permute(0, 2, 1))
    Fr = [512, 128, 32, 8]
    for i in range(4):
        C = [48, 96, 192, 384][i]
        t, pf = rs(f"yt{i}"); R = taps[f"tenc{i}"].shape[-1]
        report(f"tencoder.{i} dconv_out", t[:, pf:pf + R], taps[f"htdemucs.tencoder.{i}.dconv_out"].permute(0, 2, 1))
        t, pf = rs(f"tenc{i}"); report(f"tenc{i}", t[:, pf:pf + R], taps[f"tenc{i}"].permute(0, 2, 1))
        t, pf = rs(f"yf{i}")
        report(f"encoder.{i} dconv_out", t[:, pf:pf + Fr[i]].reshape(B, Tf, Fr[i], C),
               taps[f"htdemucs.encoder.{i}.dconv_out"].permute(0, 3, 2, 1))
        t, pf = rs(f"enc{i}"); report(f"enc{i}", t[:, pf:pf + Fr[i]].reshape(B, Tf, Fr[i], C), taps[f"enc{i}"].permute(0, 3, 2, 1))
    report("tokf (xf_layer4)", plan.tap("tokf").to_torch().view(B, -1, 512), taps["xf_layer4"])
    report("tokt (xf_layer4_t)", plan.tap("tokt").to_torch().view(B, -1, 512), taps["xf_layer4_t"])
    report("xenc", plan.tap("xenc").to_torch().view(B, Tf, 8, 384), taps["x_enc"].permute(0, 3, 2, 1))
    report("xtenc", plan.tap("xtenc").to_torch().view(B, -1, 384), taps["xt_enc"].permute(0, 2, 1))
    if a.P == 1:
        t, pf = rs("xc"); report("xc (x_cond)", t[:, pf:pf + 8].reshape(B, Tf, 8, 384), taps["x_cond"].permute(0, 3, 2, 1))
        t, pf = rs("xtc"); R = taps["xt_cond"].shape[-1]; report("xtc (xt_cond)", t[:, pf:pf + R], taps["xt_cond"].permute(0, 2, 1))
        for i in range(4):
            C = [192, 96, 48, 4][i]
            t, pf = rs(f"fdec{i}"); report(f"fdec{i}", t[:, pf:pf + Tf].reshape(B, Tf, Tf, C), taps[f"fdec{i}"].permute(0, 3, 2, 1))
            t, pf = rs(f"tdec{i}"); R = taps[f"tdec{i}"].shape[-1]; report(f"tdec{i}", t[:, pf:pf + R], taps[f"tdec{i}"].permute(0, 2, 1))
    report("OUT prompt 0", out[:, 0], ref)
    for p in range(1, a.P):
        refp = athtd_oracle.forward(sd, wav, embs[p])
        report(f"OUT prompt {p}", out[:, p], refp)


if __name__ == "__main__":
    main()
