"""Device-side state behind the module API: flat parameter buffer, packed weights, per-shape plans.

PyTorch is used here for device memory (caching allocator), streams and the constant tables that
demucs itself builds with torch (hann window, sinusoidal position embeddings); all compute goes
through libathtd.so.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, Tuple

import torch

from . import lib as _lib

DTYPES = {"fp32": 0, "bf16": 1}


def _ptr(t) -> int:
    return 0 if t is None else t.data_ptr()


def sin_embedding_1d(length: int, dim: int, max_period: float = 10000.0) -> torch.Tensor:
    """demucs transformer.py:create_sin_embedding (shift 0) as a [length, dim] table."""
    pos = torch.arange(length).view(-1, 1)
    half = dim // 2
    adim = torch.arange(half).view(1, -1)
    phase = pos / (max_period ** (adim / (half - 1)))
    return torch.cat([torch.cos(phase), torch.sin(phase)], dim=-1).float().contiguous()


def sin_embedding_2d(dim: int, height: int, width: int, max_period: float = 10000.0) -> torch.Tensor:
    """demucs transformer.py:create_2d_sin_embedding as a [(width height), dim] table in the
    "(t1 fr)" token order the cross-transformer uses (t1 = width = frames, fr = height)."""
    pe = torch.zeros(dim, height, width)
    d = dim // 2
    div = torch.exp(torch.arange(0.0, d, 2) * -(math.log(max_period) / d))
    pw = torch.arange(0.0, width).unsqueeze(1)
    ph = torch.arange(0.0, height).unsqueeze(1)
    pe[0:d:2] = torch.sin(pw * div).transpose(0, 1).unsqueeze(1).repeat(1, height, 1)
    pe[1:d:2] = torch.cos(pw * div).transpose(0, 1).unsqueeze(1).repeat(1, height, 1)
    pe[d::2] = torch.sin(ph * div).transpose(0, 1).unsqueeze(2).repeat(1, 1, width)
    pe[d + 1::2] = torch.cos(ph * div).transpose(0, 1).unsqueeze(2).repeat(1, 1, width)
    return pe.permute(2, 1, 0).reshape(width * height, dim).float().contiguous()


class Tap:
    """View of an intermediate device buffer of the last forward (tests / debugging)."""

    def __init__(self, ptr: int, numel: int, dtype: int, dims):
        self.ptr, self.numel, self.dtype, self.dims = ptr, numel, dtype, tuple(dims)

    def interior(self) -> torch.Tensor:
        """Row-space buffers: strip pad groups / pad rows -> [B*G2, R, C]."""
        groups, Rp, Cc, pf, G2, G2p, gpf, R = self.dims
        t = self.to_torch().view(groups // G2p, G2p, Rp, Cc)
        return t[:, gpf:gpf + G2, pf:pf + R].reshape(-1, R, Cc)

    def to_torch(self) -> torch.Tensor:
        tdt = torch.float32 if self.dtype == 0 else torch.bfloat16
        out = torch.empty(self.numel, dtype=tdt, device="cuda")
        _lib.check(_lib.load().athtd_memcpy_d2d(out.data_ptr(), self.ptr, self.numel * out.element_size(),
                                                torch.cuda.current_stream().cuda_stream), "athtd_memcpy_d2d")
        return out


class Plan:
    def __init__(self, engine: "Engine", B: int, L: int, P: int):
        lib = _lib.load()
        self.engine, self.B, self.L, self.P = engine, B, L, P
        dt = engine.dtype_code
        nbytes = lib.athtd_workspace_bytes(B, L, P, dt)
        if nbytes < 0:
            raise _lib.AthtdError(lib.athtd_last_error().decode())
        dev = engine.device
        self.workspace = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
        Tf = (L + 1023) // 1024
        St = L
        for _ in range(4):
            St = (St + 3) // 4
        self.Tf, self.Sf, self.St = Tf, 8 * Tf, St
        self.pe2d = sin_embedding_2d(512, 8, Tf).to(dev)
        self.pe1d = sin_embedding_1d(St, 512).to(dev)
        self.handle = lib.athtd_plan_create(B, L, P, dt, _ptr(engine.params), _ptr(engine.packed), _ptr(self.workspace),
                                            nbytes, _ptr(engine.tw), _ptr(engine.win), _ptr(self.pe2d), _ptr(self.pe1d))
        if not self.handle:
            raise _lib.AthtdError(lib.athtd_last_error().decode())

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                _lib.load().athtd_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def _stream(self) -> int:
        return torch.cuda.current_stream(self.engine.device).cuda_stream

    def forward(self, wav: torch.Tensor, emb: torch.Tensor, out: torch.Tensor = None) -> torch.Tensor:
        assert wav.shape == (self.B, 2, self.L) and wav.dtype == torch.float32 and wav.is_cuda and wav.is_contiguous()
        assert emb.shape == (self.B, self.P, 512) and emb.dtype == torch.float32 and emb.is_cuda and emb.is_contiguous()
        if out is None:
            out = torch.empty(self.B, self.P, 2, self.L, dtype=torch.float32, device=wav.device)
        _lib.check(_lib.load().athtd_forward(self.handle, wav.data_ptr(), emb.data_ptr(), out.data_ptr(), self._stream()),
                   "athtd_forward")
        return out

    def encode(self, wav: torch.Tensor) -> None:
        _lib.check(_lib.load().athtd_encode(self.handle, wav.data_ptr(), self._stream()), "athtd_encode")

    def decode(self, emb: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
        _lib.check(_lib.load().athtd_decode(self.handle, emb.data_ptr(), out.data_ptr(), self._stream()), "athtd_decode")
        return out

    @property
    def launches(self) -> int:
        return _lib.load().athtd_plan_launches(self.handle)

    def set_tc(self, on: bool) -> None:
        _lib.load().athtd_plan_set_tc(self.handle, 1 if on else 0)

    @property
    def tc_launches(self) -> int:
        return _lib.load().athtd_plan_tc_launches(self.handle)

    def set_fused_dconv(self, on: bool) -> None:
        _lib.load().athtd_plan_set_fused_dconv(self.handle, 1 if on else 0)

    def set_profile(self, on: bool) -> None:
        _lib.load().athtd_plan_set_profile(self.handle, 1 if on else 0)

    def get_profile(self):
        ms, gf, n = C.c_double(), C.c_double(), C.c_int()
        _lib.load().athtd_plan_get_profile(self.handle, C.byref(ms), C.byref(gf), C.byref(n))
        return ms.value, gf.value, n.value

    def tap(self, name: str) -> Tap:
        ptr, numel, dt = C.c_void_p(), C.c_long(), C.c_int()
        dims = (C.c_int * 8)()
        _lib.check(_lib.load().athtd_tap(self.handle, name.encode(), C.byref(ptr), C.byref(numel), C.byref(dt),
                                         C.byref(dims)), "athtd_tap")
        return Tap(ptr.value, numel.value, dt.value, list(dims))


class Engine:
    """Owns the flat fp32 parameter buffer, the packed GEMM-layout weights and the plan cache."""

    def __init__(self, device, dtype: str = "bf16"):
        if dtype not in DTYPES:
            raise ValueError(f"dtype must be one of {list(DTYPES)}")
        self.lib = _lib.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.AthtdError("AudioTextHTDemucs B200 path needs a CUDA (sm_100a) device; there is no CPU fallback")
        self.dtype = dtype
        self.dtype_code = DTYPES[dtype]
        self.table = _lib.param_table()
        self.params = torch.zeros(self.lib.athtd_params_total(), dtype=torch.float32, device=self.device)
        self.packed = torch.zeros(self.lib.athtd_packed_bytes(self.dtype_code), dtype=torch.uint8, device=self.device)
        k = torch.arange(4096, dtype=torch.float64)
        ang = -2.0 * math.pi * k / 4096.0
        self.tw = torch.stack([torch.cos(ang), torch.sin(ang)], dim=-1).float().contiguous().to(self.device)
        self.win = torch.hann_window(4096).float().contiguous().to(self.device)
        self.plans: Dict[Tuple[int, int, int], Plan] = {}
        self.loaded = False

    def load_params(self, named: Dict[str, torch.Tensor]) -> None:
        """Copy the live tensors (by reference key name) into the flat buffer and re-pack."""
        host = torch.zeros(self.params.numel(), dtype=torch.float32)
        for name, numel, off in self.table:
            if name not in named:
                raise KeyError(f"missing live parameter {name}")
            t = named[name].detach()
            if t.numel() != numel:
                raise ValueError(f"{name}: expected {numel} elements, got {tuple(t.shape)}")
            host[off:off + numel] = t.reshape(-1).float().cpu()
        with torch.cuda.device(self.device):
            self.params.copy_(host)
            _lib.check(self.lib.athtd_pack_weights(self.params.data_ptr(), self.packed.data_ptr(), self.dtype_code,
                                                   torch.cuda.current_stream().cuda_stream), "athtd_pack_weights")
        self.loaded = True

    def plan(self, B: int, L: int, P: int = 1) -> Plan:
        key = (B, L, P)
        if key not in self.plans:
            with torch.cuda.device(self.device):
                self.plans[key] = Plan(self, B, L, P)
        return self.plans[key]

    def gemm_kernel_name(self) -> str:
        return "gemm_tc_kernel + flash_attn_kernel (tcgen05)" if self.dtype == "bf16" else "gemm_simt_kernel (CUDA-core fp32)"

    def drop_plans(self) -> None:
        self.plans.clear()
