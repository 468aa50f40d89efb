"""gemm_tc on the transformer shapes with and without the epilogue stores (athtd_gemm_test mode 2 = no_store): how much of
each launch is the store path."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from athtd_b200 import lib as alib
lib = alib.load()
st = torch.cuda.current_stream().cuda_stream
for (M, N, K) in [(66304, 2048, 512), (66304, 1536, 512), (66304, 512, 512), (66304, 512, 2048), (2213408, 384, 384), (2213408, 192, 192), (2116800, 96, 48)]:
    A = torch.randn(M, K, device="cuda").bfloat16(); B = (torch.randn(N, K, device="cuda") / K ** 0.5).bfloat16()
    bias = torch.randn(N, device="cuda"); C = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
    res = []
    for mode in (1, 2):
        for _ in range(3): lib.athtd_gemm_test(A.data_ptr(), B.data_ptr(), bias.data_ptr(), C.data_ptr(), M, N, K, 1, mode, st)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): lib.athtd_gemm_test(A.data_ptr(), B.data_ptr(), bias.data_ptr(), C.data_ptr(), M, N, K, 1, mode, st)
        e1.record(); torch.cuda.synchronize()
        res.append(e0.elapsed_time(e1) / 10)
    print(f"M={M} N={N} K={K}: store {res[0]*1e3:.1f} us ({2*M*N*K/res[0]/1e9:.0f} TF/s)   no-store {res[1]*1e3:.1f} us ({2*M*N*K/res[1]/1e9:.0f} TF/s)", flush=True)
    del A, B, C
