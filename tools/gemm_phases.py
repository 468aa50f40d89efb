"""Per-CTA phase timing of gemm_tc_kernel (clock64 stamps): 0 start, 1 after setup, 2 last TMA issued, 3 last MMA committed,
4 epilogue woke, 5 epilogue done, 6 CTA end."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from athtd_b200 import lib as alib
lib = alib.load(); st = torch.cuda.current_stream().cuda_stream
for M, N, K in [(66304, 2048, 512), (66304, 512, 512), (553352, 192, 192), (8192, 8192, 8192)]:
    A = torch.randn(M, K, device="cuda").bfloat16(); B = torch.randn(N, K, device="cuda").bfloat16()
    C = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    dbg = torch.zeros(512 * 8, dtype=torch.int64, device="cuda")
    for _ in range(2):
        alib.check(lib.athtd_gemm_test(A.data_ptr(), B.data_ptr(), dbg.data_ptr(), C.data_ptr(), M, N, K, 1, 3, st))
    torch.cuda.synchronize()
    d = dbg.view(512, 8).cpu().double()
    n = min(512, (M + 127) // 128)
    d = d[:n]
    rel = d[:, 1:7] - d[:, 0:1]
    print(f"M={M} N={N} K={K}: mean cycles since CTA start  setup {rel[:,0].mean():8.0f}  tma_done {rel[:,1].mean():8.0f}  mma_done {rel[:,2].mean():8.0f}  "
          f"epi_wake {rel[:,3].mean():8.0f}  epi_done {rel[:,4].mean():8.0f}  end {rel[:,5].mean():8.0f}")
