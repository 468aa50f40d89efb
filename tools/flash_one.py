"""One launch of the fused attention kernel on the frequency-branch self-attention shape (ncu target)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from athtd_b200 import lib as alib
lib = alib.load()
st = torch.cuda.current_stream().cuda_stream
B, Sq, Sk = 32, 2072, 2072
q = torch.randn(B, Sq, 512, device="cuda").bfloat16(); k = torch.randn(B, Sk, 512, device="cuda").bfloat16()
v = torch.randn(B, Sk, 512, device="cuda").bfloat16(); o = torch.empty(B, Sq, 512, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    alib.check(lib.athtd_attention_test(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), B, Sq, Sk, st))
torch.cuda.synchronize()
print("ok")
