"""One bf16 forward at batch B (default 32) on 6 s segments -- target command for ncu captures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import athtd_b200
from athtd_b200 import synthetic as weights
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
m = athtd_b200.AudioTextHTDemucsB200(precision="bf16")
m.load_state_dict(weights.make_state_dict(0), strict=False)
m = m.cuda().eval()
m.engine().plan(B, 264600, 1).set_graph(False)      # eager launches: every kernel shows up as its own launch
wav = 0.1 * torch.randn(B, 2, 264600, device="cuda")
emb = torch.nn.functional.normalize(torch.randn(B, 512, device="cuda"), dim=-1)
for _ in range(reps):          # warm-up (weight packing, plan creation)
    out = m(wav, emb)
torch.cuda.synchronize()
torch.cuda.profiler.start()    # ncu --profile-from-start off captures exactly one forward
out = m(wav, emb)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok", tuple(out.shape), float(out.abs().mean()))
