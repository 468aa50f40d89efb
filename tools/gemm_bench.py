"""Micro-benchmark of the tcgen05 GEMM kernel on the shapes of the path (CUDA events, L2-cold via rotating buffers)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from athtd_b200 import lib as alib

SHAPES = [  # (M, N, K, note)
    (553352, 384, 384, "FreqDecoder L1 convT (B=8)"),
    (553352, 192, 192, "FreqDecoder L2 convT (B=8)"),
    (1085696, 96, 48, "enc0 rewrite (B=8)"),
    (1085696, 16, 144, "enc0 dconv k3 (B=8)"),
    (1085696, 96, 16, "enc0 dconv expand (B=8)"),
    (66304, 2048, 512, "xf linear1 (B=32)"),
    (66304, 512, 2048, "xf linear2 (B=32)"),
    (66304, 1536, 512, "xf in_proj (B=32)"),
    (66304, 512, 512, "xf out_proj (B=32)"),
    (265216, 192, 768, "enc2 conv (B=32)"),
    (8192, 8192, 8192, "square 8192"),
]
lib = alib.load()
st = torch.cuda.current_stream().cuda_stream
for M, N, K, note in SHAPES:
    A = torch.randn(M, K, device="cuda").bfloat16()
    B = (torch.randn(N, K, device="cuda") / K ** 0.5).bfloat16()
    bias = torch.zeros(N, device="cuda")
    C = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    for mode in (1, 2):
        for _ in range(3):
            alib.check(lib.athtd_gemm_test(A.data_ptr(), B.data_ptr(), bias.data_ptr(), C.data_ptr(), M, N, K, 1, mode, st))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        it = 10
        for _ in range(it):
            alib.check(lib.athtd_gemm_test(A.data_ptr(), B.data_ptr(), bias.data_ptr(), C.data_ptr(), M, N, K, 1, mode, st))
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / it
        fl = 2.0 * M * N * K
        by = 2.0 * (M * K + N * K + (M * N if mode == 1 else 0))
        print(f"{note:30s} M={M:8d} N={N:5d} K={K:5d} mode={'store' if mode == 1 else 'nostore'}: {us:9.1f} us  {fl / us * 1e-6:8.1f} TFLOP/s  {by / us * 1e-3:8.1f} GB/s")
    if (M, N, K) == (8192, 8192, 8192):
        Af, Bf = A, B
        for _ in range(3): torch.matmul(Af, Bf.t())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): torch.matmul(Af, Bf.t())
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / 10
        print(f"   cuBLAS reference for the same shape: {us:9.1f} us  {2.0 * M * N * K / us * 1e-6:8.1f} TFLOP/s")
