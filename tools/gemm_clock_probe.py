"""SM clock under a sustained run of one gemm_tc shape vs torch.matmul (cuBLAS) of the same shape: the GEMM-heavy phases of the
forward are power-capped, so TFLOP/s per wall second and per SM clock differ.  Prints median clock, TFLOP/s and the fraction of
the per-clock tensor peak (8192 dense bf16 flop / clk / SM x 148 SMs)."""
import os, subprocess, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from athtd_b200 import lib as alib
lib = alib.load()
st = torch.cuda.current_stream().cuda_stream

def sample(stop, rows):
    p = subprocess.Popen(["nvidia-smi", "--id=0", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "50"],
                         stdout=subprocess.PIPE, text=True)
    for line in p.stdout:
        rows.append(line.strip())
        if stop.is_set():
            break
    p.terminate()

for (M, N, K) in [(66304, 2048, 512), (66304, 512, 2048), (8192, 8192, 8192)]:
    A = torch.randn(M, K, device="cuda").bfloat16(); B = (torch.randn(N, K, device="cuda") / K ** 0.5).bfloat16()
    bias = torch.zeros(N, device="cuda"); C = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
    for name in ("gemm_tc", "cublas"):
        def run():
            if name == "gemm_tc": lib.athtd_gemm_test(A.data_ptr(), B.data_ptr(), bias.data_ptr(), C.data_ptr(), M, N, K, 1, 1, st)
            else: torch.matmul(A, B.t(), out=C)
        for _ in range(5): run()
        torch.cuda.synchronize()
        stop, rows = threading.Event(), []
        th = threading.Thread(target=sample, args=(stop, rows), daemon=True); th.start()
        time.sleep(0.3)
        n = 0; t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        while time.perf_counter() - t0 < 2.5:
            for _ in range(50): run()
            n += 50
            torch.cuda.synchronize()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        stop.set(); time.sleep(0.1)
        clk = sorted(int(r.split(",")[0]) for r in rows[4:] if r and r.split(",")[0].strip().isdigit())
        pw = sorted(float(r.split(",")[1]) for r in rows[4:] if "," in r)
        mhz = clk[len(clk) // 2] if clk else 0
        tf = 2.0 * M * N * K / ms / 1e9
        peak = 8192 * 148 * mhz * 1e6 / 1e12 if mhz else float("nan")
        print(f"{name:8s} M={M} N={N} K={K}: {ms*1e3:8.1f} us  {tf:7.0f} TFLOP/s  median SM clock {mhz} MHz  power {pw[len(pw)//2] if pw else 0:.0f} W  "
              f"per-clock peak {peak:.0f} TFLOP/s  fraction {tf/peak:.2f}", flush=True)
