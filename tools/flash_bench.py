"""Timing of the fused attention kernel on the path's shapes (B = 32): python tools/flash_bench.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from athtd_b200 import lib as alib
lib = alib.load()
st = torch.cuda.current_stream().cuda_stream
B = 32
import sys as _s
# every argument is one athtd_attention_set_poly() flag word (e.g. 4 0x204 6 0x104): all variants are timed in ONE process / call
flags = [int(a, 0) for a in _s.argv[1:]] or [4]
shapes = [(2072, 2072), (1034, 1034), (2072, 1034), (1034, 2072)]
bufs = {}
for (Sq, Sk) in shapes:
    q = torch.randn(B, Sq, 512, device="cuda").bfloat16(); k = torch.randn(B, Sk, 512, device="cuda").bfloat16()
    v = torch.randn(B, Sk, 512, device="cuda").bfloat16(); o = torch.empty(B, Sq, 512, device="cuda", dtype=torch.bfloat16)
    bufs[(Sq, Sk)] = (q, k, v, o)
for rep in range(2):
    for fl in flags:
        lib.athtd_attention_set_poly(fl)
        tot = 0.0
        for (Sq, Sk) in shapes:
            q, k, v, o = bufs[(Sq, Sk)]
            for _ in range(3): lib.athtd_attention_test(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), B, Sq, Sk, st)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10): lib.athtd_attention_test(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), B, Sq, Sk, st)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            tot += ms
            print(f"flags={fl:#x} Sq={Sq} Sk={Sk}: {ms*1e3:.1f} us  {4*B*8*Sq*Sk*64/ms/1e9:.0f} TFLOP/s", flush=True)
        print(f"flags={fl:#x} all four shapes: {tot*1e3:.1f} us", flush=True)
lib.athtd_attention_set_poly(4)
