"""A/B of process-wide tuning flags inside ONE process: python tools/ab_flags.py 256 1048832 ...  (athtd_set_tc_tuning values).
Times the BASELINE config-3 step (240 s track, batch 32) for each setting, alternating, three rounds."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import athtd_b200
from athtd_b200 import synthetic
lib = athtd_b200.load_library()
flags = [int(a) for a in sys.argv[1:]] or [256]
m = athtd_b200.AudioTextHTDemucsB200(precision="bf16")
m.load_state_dict(synthetic.make_state_dict(0), strict=False)
sep = athtd_b200.B200SeparationModel(m, "cuda", 6.0, 1.5, batch=32, use_graph=False)
track = synthetic.make_track(240.0).cuda()
emb = synthetic.make_prompt_embeddings(1).cuda()
n = len(athtd_b200.segment_plan(track.shape[-1]).starts)
for rnd in range(3):
    for f in flags:
        lib.athtd_set_tc_tuning(f)
        for _ in range(2):
            sep.separate_span(track, emb, (0, n))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            sep.separate_span(track, emb, (0, n))
        e1.record(); torch.cuda.synchronize()
        print(f"round {rnd} flags 0x{f:x}: {e0.elapsed_time(e1) / 10:.3f} ms per step", flush=True)
