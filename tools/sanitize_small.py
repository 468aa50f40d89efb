"""Small end-to-end run for compute-sanitizer (memcheck / racecheck, one tool per call):
    compute-sanitizer --tool memcheck python tools/sanitize_small.py
bf16 and fp32 forward at B = 2, 12 000 samples, 2 prompts (every kernel of the path incl. the fused inverse STFT and the tail-batch
plan reuse), the chunk loop with a ragged tail on a 0.5 s segment plan, the host-staged path, load_audio."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import athtd_b200
from athtd_b200 import synthetic
sd = synthetic.make_state_dict(0)
for prec in ("bf16", "fp32"):
    m = athtd_b200.AudioTextHTDemucsB200(precision=prec)
    m.load_state_dict(sd, strict=False)
    m = m.cuda().eval()
    wav, emb = synthetic.make_inputs(5, 2, 12000)
    embs = torch.stack([emb, synthetic.make_inputs(6, 2, 4096)[1]], dim=1).cuda()
    out = m.separate_batch(wav.cuda(), embs)
    out1 = m.separate_batch(wav[:1].cuda(), embs[:1])          # smaller batch in the same workspace
    sep = athtd_b200.B200SeparationModel(m, "cuda", segment_seconds=0.5, overlap_seconds=0.125, batch=3)
    T = 22050 * 2 + 3000
    mix = synthetic.make_inputs(7, 1, T)[0][0]
    a, _ = sep.separate_many(mix, embs[0])
    a2, _ = sep.separate_many(mix, embs[0])                    # second call: graph capture + replay
    host = mix.contiguous().pin_memory()
    oh = torch.empty(2, 2, T).pin_memory()
    n = len(athtd_b200.segment_plan(T, 0.5, 0.125).starts)
    sep.separate_span_host(host, embs[0], (0, n), oh)
    torch.cuda.synchronize()
    print(prec, float(out.abs().mean()), float(out1.abs().mean()), float((a - a2).abs().max()), float((oh.cuda() - a).abs().max()))
y, sr = athtd_b200.prepare_mixture(torch.randn(1, 30000), 48000)
torch.cuda.synchronize()
print("load_audio", tuple(y.shape), sr)
print("ok")
