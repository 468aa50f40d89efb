"""SM clock / board power while the batch-32 forward runs back to back for ~3 s (nvidia-smi sampled every 20 ms)."""
import os, subprocess, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import athtd_b200
from oracle import weights

dev = torch.device("cuda", 0)
model = athtd_b200.AudioTextHTDemucsB200(precision="bf16")
model.load_state_dict(weights.make_state_dict(0), strict=False)
sep = athtd_b200.B200SeparationModel(model, dev, 6.0, 1.5, batch=32)
T = 240 * 44100
track = (0.1 * torch.randn(2, T, generator=torch.Generator().manual_seed(4))).to(dev)
_, emb = weights.make_inputs(2, 1, 4096)
emb = emb.to(dev)
n = len(athtd_b200.segment_plan(T).starts)
for _ in range(3): sep.separate_span(track, emb, (0, n))
torch.cuda.synchronize()
rows, stop = [], threading.Event()
def sample():
    p = subprocess.Popen(["nvidia-smi", "--id=0", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "20"],
                         stdout=subprocess.PIPE, text=True)
    for line in p.stdout:
        rows.append(line.strip())
        if stop.is_set(): break
    p.terminate()
th = threading.Thread(target=sample, daemon=True); th.start()
time.sleep(0.3)
t0 = time.perf_counter(); k = 0
while time.perf_counter() - t0 < 3.0:
    sep.separate_span(track, emb, (0, n)); k += 1
    torch.cuda.synchronize()
dt = time.perf_counter() - t0
stop.set(); time.sleep(0.1)
clk = sorted(int(r.split(",")[0]) for r in rows[10:] if r.split(",")[0].strip().isdigit())
pw = sorted(float(r.split(",")[1]) for r in rows[10:] if "," in r)
q = lambda a, f: a[int(f * (len(a) - 1))] if a else None
print(f"forward loop: {k} tracks in {dt:.2f} s = {240 * k / dt:.0f} x realtime; SM clock MHz min/p10/median/p90/max = "
      f"{q(clk,0)}/{q(clk,.1)}/{q(clk,.5)}/{q(clk,.9)}/{q(clk,1)}; power W p10/median/p90/max = {q(pw,.1)}/{q(pw,.5)}/{q(pw,.9)}/{q(pw,1)}; samples {len(clk)}")
