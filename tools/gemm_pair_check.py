"""Correctness + timing of the CTA-pair (cta_group::2) GEMM path against torch.matmul on shapes that select it
(N % 256 == 0, M >= 2 * SMs * 128), selected with the tuning hook athtd_set_tc_tuning(256 | 0x40000).
Run under `timeout`: a protocol bug shows up as a hang."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from athtd_b200 import lib as alib
lib = alib.load()
lib.athtd_set_tc_tuning(256 | 0x40000)
st = torch.cuda.current_stream().cuda_stream
ok = True
for (M, N, K) in [(40000, 256, 64), (40000, 512, 512), (66304, 2048, 512), (66304, 512, 2048), (37889, 1536, 512)]:
    g = torch.Generator().manual_seed(M + N + K)
    A = torch.randn(M, K, generator=g).bfloat16().cuda()
    B = (torch.randn(N, K, generator=g) / K ** 0.5).bfloat16().cuda()
    bias = torch.randn(N, generator=g).cuda()
    C = torch.full((M, N), float("nan"), dtype=torch.bfloat16, device="cuda")
    alib.check(lib.athtd_gemm_test(A.data_ptr(), B.data_ptr(), bias.data_ptr(), C.data_ptr(), M, N, K, 1, 1, st))
    torch.cuda.synchronize()
    ref = (A.float() @ B.float().t() + bias)
    err = (C.float() - ref).abs().max().item()
    fin = bool(torch.isfinite(C.float()).all())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        lib.athtd_gemm_test(A.data_ptr(), B.data_ptr(), bias.data_ptr(), C.data_ptr(), M, N, K, 1, 1, st)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    good = fin and err < 6e-3 * max(1.0, ref.abs().max().item())
    ok = ok and good
    print(f"M={M} N={N} K={K}: max-abs err {err:.3e} finite={fin} {'OK' if good else 'FAIL'}  {ms*1e3:.1f} us  {2*M*N*K/ms/1e9:.0f} TFLOP/s", flush=True)
sys.exit(0 if ok else 1)
