"""Device-side state behind the module API: flat parameter buffer, packed weights, per-shape plans.

PyTorch is used here for device memory (caching allocator), streams and the constant tables that
demucs itself builds with torch (hann window, sinusoidal position embeddings); all compute goes
through libathtd.so.
"""
from __future__ import annotations

import ctypes as C
import math
from collections import OrderedDict
from typing import Optional, Tuple

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

    def __init__(self, ptr: int, numel: int, dtype: int, dims, device):
        self.ptr, self.numel, self.dtype, self.dims, self.device = ptr, numel, dtype, tuple(dims), device

    def interior(self) -> torch.Tensor:
        """Row-space buffers: strip pad groups / pad rows -> [B*G2, R, C]."""
        groups, Rp, Cc, pf, G2, G2p, gpf, R = self.dims
        t = self.to_torch().view(groups // G2p, G2p, Rp, Cc)
        return t[:, gpf:gpf + G2, pf:pf + R].reshape(-1, R, Cc)

    def to_torch(self) -> torch.Tensor:
        tdt = torch.float32 if self.dtype == 0 else torch.bfloat16
        out = torch.empty(self.numel, dtype=tdt, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().athtd_memcpy_d2d(out.data_ptr(), self.ptr, self.numel * out.element_size(),
                                                    torch.cuda.current_stream(self.device).cuda_stream), "athtd_memcpy_d2d")
        return out


class Plan:
    """One (segment length L, prompts P) launch sequence with a workspace laid out for ``cap`` segments; any batch
    B <= cap runs in it (``set_batch`` / the leading dimension of ``wav``), so a track's tail batch shares the
    workspace of the full batches."""

    def __init__(self, engine: "Engine", cap: int, L: int, P: int):
        lib = _lib.load()
        self.engine, self.cap, self.B, self.L, self.P = engine, cap, cap, L, P
        B = cap
        dt = engine.dtype_code
        nbytes = lib.athtd_workspace_bytes(B, L, P, dt)
        if nbytes < 0:
            raise _lib.AthtdError(lib.athtd_last_error().decode())
        dev = engine.device
        self.nbytes = nbytes
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

    def set_batch(self, B: int) -> "Plan":
        if B != self.B:
            _lib.check(_lib.load().athtd_plan_set_batch(self.handle, B), "athtd_plan_set_batch")
            self.B = B
        return self

    def _check_wav(self, wav: torch.Tensor) -> None:
        if wav.dim() != 3 or wav.shape[1] != 2 or wav.shape[2] != self.L or not 1 <= wav.shape[0] <= self.cap:
            raise ValueError(f"wav must be [B<={self.cap}, 2, {self.L}], got {tuple(wav.shape)}")
        if wav.dtype != torch.float32 or wav.device != self.engine.device or not wav.is_contiguous():
            raise ValueError("wav must be a contiguous float32 tensor on the engine's device")

    def _check_emb(self, emb: torch.Tensor) -> None:
        if emb.shape != (self.B, self.P, 512) or emb.dtype != torch.float32 or emb.device != self.engine.device or not emb.is_contiguous():
            raise ValueError(f"emb must be a contiguous float32 [{self.B}, {self.P}, 512] tensor on the engine's device")

    def forward(self, wav: torch.Tensor, emb: torch.Tensor, out: torch.Tensor = None) -> torch.Tensor:
        self._check_wav(wav)
        self.set_batch(wav.shape[0])
        self._check_emb(emb)
        if out is None:
            out = torch.empty(self.B, self.P, 2, self.L, dtype=torch.float32, device=wav.device)
        with torch.cuda.device(self.engine.device):
            _lib.check(_lib.load().athtd_forward(self.handle, wav.data_ptr(), emb.data_ptr(), out.data_ptr(), self._stream()),
                       "athtd_forward")
        return out

    def encode(self, wav: torch.Tensor) -> None:
        """Prompt-independent half (ATHTDemucs_v2.py:261-279); the encoder state stays in the workspace."""
        self._check_wav(wav)
        self.set_batch(wav.shape[0])
        with torch.cuda.device(self.engine.device):
            _lib.check(_lib.load().athtd_encode(self.handle, wav.data_ptr(), self._stream()), "athtd_encode")

    def decode(self, emb: torch.Tensor, out: torch.Tensor = None) -> torch.Tensor:
        """Per-prompt half (:282-324) over the state of the last ``encode``; may be called any number of times."""
        self._check_emb(emb)
        if out is None:
            out = torch.empty(self.B, self.P, 2, self.L, dtype=torch.float32, device=emb.device)
        with torch.cuda.device(self.engine.device):
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

    def set_graph(self, on: bool) -> None:
        _lib.load().athtd_plan_set_graph(self.handle, 1 if on else 0)

    @property
    def graph_replays(self) -> int:
        return _lib.load().athtd_plan_graph_replays(self.handle)

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
        return Tap(ptr.value, numel.value, dt.value, list(dims), self.engine.device)


class Engine:
    """Owns the flat fp32 parameter buffer, the packed GEMM-layout weights and the plan cache."""

    def __init__(self, device, dtype: str = "bf16", max_workspace_bytes: Optional[int] = None):
        if dtype not in DTYPES:
            raise ValueError(f"dtype must be one of {list(DTYPES)}")
        self.lib = _lib.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.AthtdError("AudioTextHTDemucs B200 path needs a CUDA (sm_100a) device; there is no CPU fallback")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.dtype = dtype
        self.dtype_code = DTYPES[dtype]
        self.table = _lib.param_table()
        self.params = torch.zeros(self.lib.athtd_params_total(), dtype=torch.float32, device=self.device)
        self.packed = torch.zeros(self.lib.athtd_packed_bytes(self.dtype_code), dtype=torch.uint8, device=self.device)
        k = torch.arange(4096, dtype=torch.float64)
        ang = -2.0 * math.pi * k / 4096.0
        self.tw = torch.stack([torch.cos(ang), torch.sin(ang)], dim=-1).float().contiguous().to(self.device)
        self.win = torch.hann_window(4096).float().contiguous().to(self.device)
        # plan cache: one plan per (L, P), least recently used first; workspaces (0.35 GB per segment of capacity at 6 s)
        # are evicted beyond ``max_workspace_bytes`` (default: half of the device memory)
        self.plans: "OrderedDict[Tuple[int, ...], Plan]" = OrderedDict()
        if max_workspace_bytes is None:
            max_workspace_bytes = torch.cuda.get_device_properties(self.device).total_memory // 2
        self.max_workspace_bytes = int(max_workspace_bytes)
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

    def plan(self, B: int, L: int, P: int = 1, cap: Optional[int] = None, slot: int = 0) -> Plan:
        """Plan for segments of L samples and P prompts, set to batch B.  ``cap`` (>= B) is the batch capacity to lay the
        workspace out for when a new one is needed (a track loop passes its full batch size so that the tail batch
        reuses the workspace).  ``slot`` > 0 asks for a further workspace of the same shape (a track loop that keeps two
        batches in flight on two streams)."""
        key = (L, P) if slot == 0 else (L, P, slot)
        pl = self.plans.get(key)
        if pl is not None and pl.cap < B:
            del self.plans[key]       # outgrown: replaced below
            pl = None
        if pl is None:
            cap = max(B, cap or B)
            need = _lib.load().athtd_workspace_bytes(cap, L, P, self.dtype_code)
            if need < 0:
                raise _lib.AthtdError(_lib.load().athtd_last_error().decode())
            while self.plans and self.workspace_bytes() + need > self.max_workspace_bytes:
                self.plans.popitem(last=False)                 # evict the least recently used workspace
            with torch.cuda.device(self.device):
                pl = Plan(self, cap, L, P)
            self.plans[key] = pl
        self.plans.move_to_end(key)
        return pl.set_batch(B)

    def workspace_bytes(self) -> int:
        return sum(p.nbytes for p in self.plans.values())

    def gemm_kernel_name(self) -> str:
        return "gemm_tc_kernel + flash_attn_kernel (tcgen05)" if self.dtype == "bf16" else "gemm_simt_kernel (CUDA-core fp32)"

    def drop_plans(self) -> None:
        self.plans.clear()
