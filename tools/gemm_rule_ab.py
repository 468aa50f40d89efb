"""A/B of the tcgen05 GEMM tile-selection rules inside one batch-32 forward: per-launch CUDA-event times of every GEMM / attention
launch for each setting of athtd_set_tc_tuning (python tools/gemm_rule_ab.py 256 131328 ...)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import athtd_b200
from athtd_b200 import lib as alib, synthetic
os.environ["ATHTD_PROFILE_DUMP"] = "1"
B = 32
m = athtd_b200.AudioTextHTDemucsB200(precision="bf16")
m.load_state_dict(synthetic.make_state_dict(0), strict=False)
m = m.cuda().eval()
wav = 0.1 * torch.randn(B, 2, 264600, device="cuda")
emb = torch.nn.functional.normalize(torch.randn(B, 512, device="cuda"), dim=-1)
plan = m.engine().plan(B, 264600, 1)
plan.set_graph(False)
for flags in [int(a) for a in sys.argv[1:]] or [256]:
    alib.load().athtd_set_tc_tuning(flags)
    for _ in range(2):
        m(wav, emb)
    torch.cuda.synchronize()
    plan.set_profile(True)
    m(wav, emb)
    torch.cuda.synchronize()
    print(f"=== tuning flags {flags} (0x{flags:x})", file=sys.stderr, flush=True)
    ms, gf, n = plan.get_profile()
    plan.set_profile(False)
    print(f"flags 0x{flags:x}: {n} launches, {ms:.3f} ms, {gf / ms:.1f} TFLOP/s", flush=True)
