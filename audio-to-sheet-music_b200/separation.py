"""Track-level separation: segment plan, batched forward, weighted overlap-add.

Mirrors the reference's plugin interface ``SeparationModel`` (/root/reference/benchmark.py:81-115)
and replaces ``OurModel._chunked_inference`` / ``separate`` / ``separate_all``
(benchmark.py:155-215; same loop in app.py:129-178).  The python ``while`` loop over 6 s chunks
(batch 1, one prompt per pass) becomes: gather all chunks of a span on the device, run them through
the model in batches of ``batch`` segments with ALL prompts decoded from one shared encoder pass,
and overlap-add with one gather kernel that reproduces the reference's fp32 operation order.
"""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Dict, List, NamedTuple, Optional, Sequence

import torch

from . import lib as _lib
from .model import AudioTextHTDemucsB200

SAMPLE_RATE = 44100
STEMS = ["drums", "bass", "other", "vocals"]      # benchmark.py:58


class SeparationModel(ABC):
    """Same contract as the reference's plugin base class, benchmark.py:81-115: ``separate`` one stem,
    ``separate_all`` stems, ``name``."""

    @abstractmethod
    def separate(self, mixture: torch.Tensor, stem_name: str) -> torch.Tensor:
        """mixture (C, T) -> the named stem (C, T)."""

    @abstractmethod
    def separate_all(self, mixture: torch.Tensor) -> Dict[str, torch.Tensor]:
        """mixture (C, T) -> {stem name: (C, T)}."""

    @property
    @abstractmethod
    def name(self) -> str:
        """Model name for display."""


class SegmentPlan(NamedTuple):
    T: int
    chunk_len: int
    stride: int
    starts: List[int]
    ends: List[int]
    actual_len: List[int]
    fade_len: List[int]
    flags: List[int]          # bit0 fade-in, bit1 fade-out


def segment_plan(T: int, segment_seconds: float = 6.0, overlap_seconds: float = 1.5,
                 sample_rate: int = SAMPLE_RATE) -> SegmentPlan:
    """Integer arithmetic of benchmark.py:158-198 (loop bound ``start < T``, fade_len =
    min(overlap, actual_len // 2), fade-in iff start > 0, fade-out iff end < T)."""
    chunk_len = int(sample_rate * segment_seconds)
    overlap = int(overlap_seconds * sample_rate)
    stride = chunk_len - overlap
    if stride <= 0 or 2 * stride <= chunk_len:
        raise ValueError("overlap must be smaller than half a segment (at most two chunks may overlap a sample)")
    starts, ends, actual, fades, flags = [], [], [], [], []
    n = (T + stride - 1) // stride if T > 0 else 0
    for k in range(n):
        s = k * stride
        e = min(s + chunk_len, T)
        a = e - s
        f = min(overlap, a // 2)
        starts.append(s); ends.append(e); actual.append(a); fades.append(f)
        flags.append((1 if (s > 0 and f > 0) else 0) | (2 if (e < T and f > 0) else 0))
    return SegmentPlan(T, chunk_len, stride, starts, ends, actual, fades, flags)


class OlaTables:
    """Device-side description of a SegmentPlan for athtd_chunk_ola.  The fade ramps are produced by
    ``torch.linspace`` on the CPU, exactly like the oracle's restatement of benchmark.py:187-192."""

    def __init__(self, plan: SegmentPlan, device):
        self.plan = plan
        lens = sorted(set(f for f in plan.fade_len if f > 0))
        offs, up, down, o = {}, [], [], 0
        for f in lens:
            offs[f] = o
            up.append(torch.linspace(0, 1, f)); down.append(torch.linspace(1, 0, f)); o += f
        z = torch.zeros(1)
        self.ramp_up = (torch.cat(up) if up else z).float().to(device)
        self.ramp_down = (torch.cat(down) if down else z).float().to(device)
        self.ramp_off = torch.tensor([offs.get(f, 0) for f in plan.fade_len], dtype=torch.int32, device=device)
        self.starts = torch.tensor(plan.starts, dtype=torch.int64, device=device)
        self.actual = torch.tensor(plan.actual_len, dtype=torch.int32, device=device)
        self.fade = torch.tensor(plan.fade_len, dtype=torch.int32, device=device)
        self.flags = torch.tensor(plan.flags, dtype=torch.int32, device=device)


def gather_chunks(track: torch.Tensor, tables: OlaTables, k0: int, k1: int) -> torch.Tensor:
    """track [2,T] cuda -> segments [k1-k0, 2, chunk_len] (tail zero-padded, benchmark.py:170-172)."""
    p = tables.plan
    C, T = track.shape
    segs = torch.empty(k1 - k0, C, p.chunk_len, dtype=torch.float32, device=track.device)
    st = torch.cuda.current_stream(track.device).cuda_stream
    _lib.check(_lib.load().athtd_gather_chunks(track.data_ptr(), T, C, tables.starts[k0:].data_ptr(), k1 - k0, p.chunk_len,
                                               segs.data_ptr(), st), "athtd_gather_chunks")
    return segs


def chunk_ola(seg_out: torch.Tensor, seg_stride: int, k_base: int, tables: OlaTables, t_begin: int, t_end: int,
              out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Weighted overlap-add of per-chunk outputs (benchmark.py:185-202) for track samples
    [t_begin, t_end).  seg_out holds chunk k at element offset (k - k_base) * seg_stride as [2, chunk_len]."""
    p = tables.plan
    if out is None:
        out = torch.empty(2, t_end - t_begin, dtype=torch.float32, device=seg_out.device)
    st = torch.cuda.current_stream(seg_out.device).cuda_stream
    _lib.check(_lib.load().athtd_chunk_ola(seg_out.data_ptr(), seg_stride, k_base, p.chunk_len, tables.starts.data_ptr(),
                                           tables.actual.data_ptr(), tables.fade.data_ptr(), tables.flags.data_ptr(),
                                           len(p.starts), p.stride, tables.ramp_up.data_ptr(), tables.ramp_down.data_ptr(),
                                           tables.ramp_off.data_ptr(), out.data_ptr(), 2, t_begin, t_end, st),
               "athtd_chunk_ola")
    return out


class B200SeparationModel(SeparationModel):
    """``OurModel`` (benchmark.py:122-215) on the B200 path.  ``model`` is an AudioTextHTDemucsB200
    whose weights were loaded the usual way (load_state_dict(ckpt["model_state_dict"], strict=False)).

    ``batch`` segments go through the network per launch sequence; ``overlap_batches`` (default) keeps TWO such batches in flight
    on two streams with one workspace each (11 GB per workspace at batch 32), so that the ramp / tail of one batch's kernels is
    filled by the other's (measured -1 % per track); the results are bit-identical to one batch at a time."""

    def __init__(self, model: AudioTextHTDemucsB200, device: str = "cuda", segment_seconds: float = 6.0,
                 overlap_seconds: float = 1.5, batch: int = 32, use_graph: bool = True, overlap_batches: bool = True):
        self.model = model.to(device).eval()
        self.use_graph = use_graph
        self.device = torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.segment_seconds = segment_seconds
        self.overlap = overlap_seconds
        self.batch = batch
        self.last_launches = 0
        self._last_seg_out = None
        self._tables_key = None
        self._streams = None
        self._side = None
        self.overlap_batches = overlap_batches
        self._stage = {}
        self._bufs = {}
        self.host_done: Optional[torch.cuda.Event] = None

    @property
    def name(self) -> str:
        return "AudioTextHTDemucs (B200)"

    def prompt_embeddings(self, prompts: Sequence[str]) -> torch.Tensor:
        return self.model._get_clap_embeddings(list(prompts), self.device)          # [P, 512]

    # ------------------------------------------------------------------ shared pieces
    def _plan_and_tables(self, T: int):
        plan = segment_plan(T, self.segment_seconds, self.overlap, self.model.sample_rate)
        key = (T, self.segment_seconds, self.overlap, str(self.device))
        if self._tables_key != key:
            self._tables, self._tables_key = OlaTables(plan, self.device), key
        return plan, self._tables

    def _buffers(self, kind: str, nk: int, P: int, L: int, seg_rows: int):
        """Staging buffers that persist across calls of the same shape (segments, per-chunk outputs): stable device pointers
        let the plan replay its launch sequence as a CUDA graph (engine.Plan / athtd_plan_set_graph) instead of re-launching
        ~160 kernels per batch.  One entry per call kind; a different shape replaces it."""
        key = (nk, P, L, seg_rows)
        ent = self._bufs.get(kind)
        if ent is None or ent["key"] != key:
            ent = {"key": key,
                   "segs": torch.empty(seg_rows, 2, L, dtype=torch.float32, device=self.device),
                   "seg_out": torch.empty(nk + 1, P, 2, L, dtype=torch.float32, device=self.device),     # slot 0 = halo chunk k0-1
                   "emb": {}}
            self._bufs[kind] = ent
        return ent

    def release_buffers(self) -> None:
        """Free the persistent staging buffers (and with them the graph-friendly pointers)."""
        self._bufs.clear()
        self._stage.clear()

    def _forward_batch(self, segs: torch.Tensor, emb: torch.Tensor, out: torch.Tensor, emb_cache: Optional[dict] = None,
                       slot: int = 0) -> int:
        """segs [b, 2, L] -> out [b, P, 2, L] through the (L, P) plan laid out for ``self.batch`` segments (``slot``: which of the
        workspaces of that shape, for batches in flight at the same time)."""
        b, _, L = segs.shape
        P = emb.shape[0]
        fplan = self.model.engine(self.device).plan(b, L, P, cap=self.batch, slot=slot)
        if not self.use_graph:
            fplan.set_graph(False)
        if emb_cache is not None:
            e = emb_cache.get((b, slot))
            if e is None:
                e = emb_cache[(b, slot)] = torch.empty(b, P, 512, dtype=torch.float32, device=self.device)
            e.copy_(emb.unsqueeze(0).expand(b, P, 512))
        else:
            e = emb.unsqueeze(0).expand(b, P, 512).contiguous()
        fplan.forward(segs, e, out)
        return fplan.launches

    # ------------------------------------------------------------------ device-resident span
    @torch.no_grad()
    def separate_span(self, track: torch.Tensor, emb: torch.Tensor, span, halo_exchange=None, track_offset: int = 0,
                      track_len: Optional[int] = None, halo_in: Optional[torch.Tensor] = None):
        """Separate the chunk span ``(k0, k1)`` of a track for all prompts.

        track : [2, Tpart] float32 on the device; sample 0 of it is global track sample ``track_offset``
                (it must cover every chunk of the span); ``track_len`` = global track length T.
        emb   : [P, 512].   Returns [P, 2, t_end - t_begin] for global samples [starts[k0], starts[k1]) (to T at the end).
        halo_exchange(last_chunk_out[P,2,L]) -> left neighbour's last chunk output (or None): the one
        seam exchange a multi-GPU run needs (each sample has at most two contributing chunks).  An empty span
        (k0 == k1: more ranks than chunks) returns [P, 2, 0]; build ``halo_exchange`` with the list of spans so that
        empty ranks are skipped by their neighbours."""
        T = track.shape[-1] + track_offset if track_len is None else track_len
        P = emb.shape[0]
        plan, tables = self._plan_and_tables(T)
        n = len(plan.starts)
        k0, k1 = span
        L = plan.chunk_len
        nk = k1 - k0
        if nk <= 0:
            if halo_exchange is not None:
                halo_exchange(torch.zeros(P, 2, L, dtype=torch.float32, device=self.device))
            self.last_launches = 0
            self._last_seg_out = None
            return torch.empty(P, 2, 0, dtype=torch.float32, device=self.device)
        bufs = self._buffers("span", nk, P, L, nk)
        seg_out, segs = bufs["seg_out"], bufs["segs"]
        local_starts = (tables.starts[k0:k1] - track_offset).contiguous()
        st = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(_lib.load().athtd_gather_chunks(track.data_ptr(), track.shape[-1], 2, local_starts.data_ptr(), nk, L,
                                                   segs.data_ptr(), st), "athtd_gather_chunks")
        launches = 1
        if self.overlap_batches and nk > self.batch:
            # two batches in flight: consecutive batches alternate between two workspaces on two side streams, so that the ramp /
            # tail of one batch's kernels (and the short grids of the deep levels) are filled by the other batch's CTAs.  Batches
            # read / write disjoint slices of the staging buffers; the overlap-add below runs after both streams have joined.
            cur = torch.cuda.current_stream(self.device)
            if self._side is None:
                self._side = (torch.cuda.Stream(self.device), torch.cuda.Stream(self.device))
            for s in self._side:
                s.wait_stream(cur)
            for i, b0 in enumerate(range(0, nk, self.batch)):
                b1 = min(b0 + self.batch, nk)
                with torch.cuda.stream(self._side[i & 1]):
                    launches += self._forward_batch(segs[b0:b1], emb, seg_out[1 + b0:1 + b1], bufs["emb"], slot=i & 1)
            for s in self._side:
                cur.wait_stream(s)
        else:
            for b0 in range(0, nk, self.batch):
                b1 = min(b0 + self.batch, nk)
                launches += self._forward_batch(segs[b0:b1], emb, seg_out[1 + b0:1 + b1], bufs["emb"])
        if halo_exchange is not None:
            halo_in = halo_exchange(seg_out[nk])
        if k0 > 0:
            if halo_in is None:
                raise ValueError("span starts inside the track: the left neighbour's last chunk output is required")
            seg_out[0].copy_(halo_in)
        t_begin = plan.starts[k0]
        t_end = plan.T if k1 == n else plan.starts[k1]
        out = torch.empty(P, 2, t_end - t_begin, dtype=torch.float32, device=self.device)
        for p in range(P):
            chunk_ola(seg_out[:, p], P * 2 * L, k0 - 1, tables, t_begin, t_end, out[p])
        self.last_launches = launches + P
        self._last_seg_out = seg_out
        return out

    # ------------------------------------------------------------------ host-staged span (row f-1)
    @torch.no_grad()
    def separate_span_host(self, track_host: torch.Tensor, emb: torch.Tensor, span, out_host: torch.Tensor,
                           halo_exchange=None, halo_in: Optional[torch.Tensor] = None, wait: bool = True):
        """Host-staged variant of ``separate_span`` (the step either side of the path: app.py:230-231,
        benchmark.py:207/211 move the whole track with ``.to(device)`` / ``.cpu()`` around the loop).

        track_host : [2, T] float32 in PINNED host memory (the whole track); out_host : [P, 2, t_end - t_begin] pinned.
        Inputs are uploaded batch by batch on an upload stream while the previous batch computes; every batch is
        overlap-added as soon as its chunks (and the chunk before them) are done and its output range is downloaded on a
        separate download stream while the next batch computes.  The device staging buffers are double-buffered per call,
        so with ``wait=False`` the next call's uploads overlap this call's compute and downloads (a service separating
        tracks back to back): the caller then waits on ``self.host_done`` (a CUDA event recorded after the last download)
        before reading ``out_host``.  ``wait=True`` (default) returns after the last download has completed.
        Results are bit-identical to ``separate_span`` (same kernels, same order of operations per sample)."""
        T = track_host.shape[-1]
        P = emb.shape[0]
        plan, tables = self._plan_and_tables(T)
        n = len(plan.starts)
        k0, k1 = span
        L = plan.chunk_len
        nk = k1 - k0
        dev = self.device
        comp = torch.cuda.current_stream(dev)
        if self._streams is None:
            self._streams = (torch.cuda.Stream(dev), torch.cuda.Stream(dev))
        up, down = self._streams
        if nk <= 0:
            if halo_exchange is not None:
                halo_exchange(torch.zeros(P, 2, L, dtype=torch.float32, device=dev))
            self.last_launches = 0
            self._last_seg_out = None
            self.host_done = torch.cuda.Event()
            self.host_done.record(down)
            return out_host
        in_lo, in_hi = plan.starts[k0], min(T, plan.starts[k1 - 1] + L)
        t_begin = plan.starts[k0]
        t_end = plan.T if k1 == n else plan.starts[k1]
        # double-buffered device staging of the input range: buffer `par` was last read by the gathers of the call before the
        # previous one, so this call's uploads only wait for THAT (not for the previous call's compute)
        skey = (in_hi - in_lo, str(dev))
        stage = self._stage.get(skey)
        if stage is None:
            if self._stage:
                torch.cuda.synchronize(dev)        # the old staging buffers may still be in use on the copy streams
            self._stage.clear()
            stage = {"buf": [torch.empty(2, in_hi - in_lo, dtype=torch.float32, device=dev) for _ in range(2)],
                     "free": [None, None], "par": 0}
            self._stage[skey] = stage
            up.wait_stream(comp)                   # the buffers were allocated on the compute stream
        par = stage["par"]
        stage["par"] = 1 - par
        track = stage["buf"][par]
        for ev in stage["free"][par] or ():
            up.wait_event(ev)
        batches = [(b0, min(b0 + self.batch, nk)) for b0 in range(0, nk, self.batch)]
        rows = min(self.batch, nk)
        overlap = self.overlap_batches and len(batches) > 1      # two batches in flight (see separate_span)
        bufs = self._buffers("host", nk, P, L, 2 * rows if overlap else rows)
        seg_out, segs = bufs["seg_out"], bufs["segs"]
        local_starts = (tables.starts[k0:k1] - in_lo).contiguous()
        if overlap:
            if self._side is None:
                self._side = (torch.cuda.Stream(dev), torch.cuda.Stream(dev))
            for sd in self._side:
                sd.wait_stream(comp)               # the previous call's overlap-add still reads seg_out on the compute stream

        def upload(i, done_to):
            b0, b1 = batches[i]
            hi = min(T, plan.starts[k0 + b1 - 1] + L)
            with torch.cuda.stream(up):
                for c in range(2):
                    track[c, done_to - in_lo:hi - in_lo].copy_(track_host[c, done_to:hi], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(up)
            return ev, hi

        launches = 0
        up_ev, up_to = upload(0, in_lo)
        pending_first = None
        gather_done = []
        for i, (b0, b1) in enumerate(batches):
            slot = (i & 1) if overlap else 0
            fs = self._side[slot] if overlap else comp          # stream of this batch's gather + forward
            fs.wait_event(up_ev)
            if i + 1 < len(batches):
                up_ev, up_to = upload(i + 1, up_to)
            sg = segs[slot * rows:slot * rows + (b1 - b0)]
            _lib.check(_lib.load().athtd_gather_chunks(track.data_ptr(), track.shape[-1], 2, local_starts[b0:].data_ptr(), b1 - b0, L,
                                                       sg.data_ptr(), fs.cuda_stream), "athtd_gather_chunks")
            if i >= len(batches) - (2 if overlap else 1):       # last gather(s) that read the staged input range
                ev = torch.cuda.Event()
                ev.record(fs)
                gather_done.append(ev)
            with torch.cuda.stream(fs):
                launches += self._forward_batch(sg, emb, seg_out[1 + b0:1 + b1], bufs["emb"], slot=slot) + 1
            if overlap:
                ev = torch.cuda.Event()
                ev.record(fs)
                comp.wait_event(ev)                              # overlap-add / halo exchange below run on the compute stream
            if i == len(batches) - 1 and halo_exchange is not None:
                halo_in = halo_exchange(seg_out[nk])
            # output samples that are complete now: [starts[k0+b0], starts[k0+b1]) (to the span end for the last batch)
            ra = plan.starts[k0 + b0]
            rb = t_end if b1 == nk else plan.starts[k0 + b1]
            if b0 == 0 and k0 > 0 and halo_in is None and halo_exchange is not None:
                pending_first = (ra, rb)             # needs the left neighbour's last chunk: finished after the exchange
                continue
            if b0 == 0 and k0 > 0:
                if halo_in is None:
                    raise ValueError("span starts inside the track: the left neighbour's last chunk output is required")
                seg_out[0].copy_(halo_in)
            self._ola_and_download(seg_out, tables, k0, P, L, ra, rb, t_begin, out_host, comp, down)
            launches += P
        if pending_first is not None:
            if halo_in is None:
                raise ValueError("span starts inside the track: the left neighbour's last chunk output is required")
            seg_out[0].copy_(halo_in)
            self._ola_and_download(seg_out, tables, k0, P, L, pending_first[0], pending_first[1], t_begin, out_host, comp, down)
            launches += P
        stage["free"][par] = gather_done
        self.host_done = torch.cuda.Event()
        self.host_done.record(down)
        if wait:
            self.host_done.synchronize()
        self.last_launches = launches
        self._last_seg_out = seg_out
        return out_host

    def _ola_and_download(self, seg_out, tables, k0, P, L, ra, rb, t_begin, out_host, comp, down):
        piece = torch.empty(P, 2, rb - ra, dtype=torch.float32, device=self.device)
        for p in range(P):
            chunk_ola(seg_out[:, p], P * 2 * L, k0 - 1, tables, ra, rb, piece[p])
        ev = torch.cuda.Event()
        ev.record(comp)
        with torch.cuda.stream(down):
            down.wait_event(ev)
            for p in range(P):
                for c in range(2):
                    out_host[p, c, ra - t_begin:rb - t_begin].copy_(piece[p, c], non_blocking=True)
        piece.record_stream(down)

    def separate_many(self, mixture: torch.Tensor, emb: torch.Tensor, span=None, halo_in: Optional[torch.Tensor] = None):
        """mixture [2,T], emb [P,512] -> ([P,2,T'], raw output [P,2,chunk_len] of the span's last chunk)."""
        mixture = mixture.to(self.device, torch.float32).contiguous()
        emb = emb.to(self.device, torch.float32).contiguous()
        if span is None:
            n = len(segment_plan(mixture.shape[-1], self.segment_seconds, self.overlap, self.model.sample_rate).starts)
            span = (0, n)
        out = self.separate_span(mixture, emb, span, halo_in=halo_in)
        # (a copy: the per-chunk outputs live in a staging buffer that the next call of the same shape overwrites)
        return out, (self._last_seg_out[-1].clone() if self._last_seg_out is not None else None)

    @torch.no_grad()
    def profile_gemms(self, track: torch.Tensor, emb: torch.Tensor, span, track_offset: int = 0, track_len: Optional[int] = None):
        """Untimed measurement pass: CUDA events around every GEMM / attention launch of one step (bench.py roofline).
        ``track`` / ``track_offset`` / ``track_len`` as for ``separate_span``."""
        eng = self.model.engine(self.device)
        L = int(self.model.sample_rate * self.segment_seconds)
        pl = eng.plan(min(self.batch, max(1, span[1] - span[0])), L, emb.shape[0], cap=self.batch)
        pl.set_profile(True)
        overlap, self.overlap_batches = self.overlap_batches, False      # one batch at a time through the profiled workspace: the
        try:                                                             # per-launch events must not see another stream's kernels
            self.separate_span(track, emb, span, track_offset=track_offset, track_len=track_len,
                               halo_in=torch.zeros(emb.shape[0], 2, L, device=self.device))
            torch.cuda.synchronize(self.device)
        finally:
            self.overlap_batches = overlap
        ms, gf, n = pl.get_profile()
        pl.set_profile(False)
        return {"kernel": eng.gemm_kernel_name(), "ms": ms, "gflop": gf, "launches": n, "tflops": gf / ms if ms > 0 else 0.0}

    # ------------------------------------------------------------------ test_inference.py loop (row f-4)
    @torch.no_grad()
    def separate_fade(self, mixture: torch.Tensor, emb: torch.Tensor, segment_seconds: Optional[float] = None,
                      overlap_seconds: float = 0.1) -> torch.Tensor:
        """The chunk loop of test_inference.py:96-141 (the reference's older inference script): stride chunk_len - overlap,
        the last chunk is NOT zero-padded (the model runs on its true length), every chunk is multiplied by a
        ``torchaudio.transforms.Fade(fade_in, fade_out, "linear")`` mask (fade-in iff start > 0, fade-out iff end < T, both of
        ``int(overlap * sr)`` samples) and added into the output; there is no weight normalisation.
        mixture [2, T], emb [P, 512] -> [P, 2, T].  Chunks of equal length share one batched launch sequence."""
        mixture = mixture.to(self.device, torch.float32).contiguous()
        emb = emb.to(self.device, torch.float32).contiguous()
        sr = self.model.sample_rate
        seg = self.segment_seconds if segment_seconds is None else segment_seconds
        T = mixture.shape[-1]
        P = emb.shape[0]
        L = int(sr * seg)
        ov = int(overlap_seconds * sr)
        stride = L - ov
        if stride <= 0 or 2 * stride <= L:
            raise ValueError("overlap must be smaller than half a segment")
        starts = list(range(0, T, stride))
        ends = [min(s + L, T) for s in starts]
        actual = [e - s for s, e in zip(starts, ends)]
        flags = [(1 if s > 0 else 0) | (2 if e < T else 0) for s, e in zip(starts, ends)]
        for a, f in zip(actual, flags):
            if f and a < ov:       # torchaudio's Fade would fail on torch.ones(negative)
                raise ValueError(f"a faded chunk of {a} samples is shorter than the fade ({ov}): the reference loop fails here too")
            if a < 4096:
                raise ValueError(f"chunk of {a} samples is shorter than one STFT window")
        n = len(starts)
        seg_out = torch.zeros(n, P, 2, L, dtype=torch.float32, device=self.device)
        st = torch.cuda.current_stream(self.device).cuda_stream
        full = [k for k in range(n) if actual[k] == L]
        for b0 in range(0, len(full), self.batch):
            ks = full[b0:b0 + self.batch]                       # consecutive chunk indices
            segs = torch.empty(len(ks), 2, L, dtype=torch.float32, device=self.device)
            st_dev = torch.tensor([starts[k] for k in ks], dtype=torch.int64, device=self.device)
            _lib.check(_lib.load().athtd_gather_chunks(mixture.data_ptr(), T, 2, st_dev.data_ptr(), len(ks), L, segs.data_ptr(), st),
                       "athtd_gather_chunks")
            self._forward_batch(segs, emb, seg_out[ks[0]:ks[0] + len(ks)])
        eng = self.model.engine(self.device)
        for k in range(n):
            if actual[k] != L:                                  # ragged tail chunk(s): the model sees the true length
                chunk = mixture[:, starts[k]:ends[k]].unsqueeze(0).contiguous()
                o = torch.empty(1, P, 2, actual[k], dtype=torch.float32, device=self.device)
                eng.plan(1, actual[k], P).forward(chunk, emb.unsqueeze(0).contiguous(), o)
                seg_out[k, :, :, :actual[k]] = o[0]
        up = torch.linspace(0, 1, ov).clamp_(0, 1) if ov > 0 else torch.zeros(1)          # Fade._fade_in (linear)
        down = (-torch.linspace(0, 1, ov) + 1).clamp_(0, 1) if ov > 0 else torch.zeros(1)  # Fade._fade_out (linear)
        dev = self.device
        t_starts = torch.tensor(starts, dtype=torch.int64, device=dev)
        t_actual = torch.tensor(actual, dtype=torch.int32, device=dev)
        t_fade = torch.full((n,), ov, dtype=torch.int32, device=dev)
        t_flags = torch.tensor(flags if ov > 0 else [0] * n, dtype=torch.int32, device=dev)
        t_off = torch.zeros(n, dtype=torch.int32, device=dev)
        up, down = up.float().to(dev), down.float().to(dev)
        out = torch.empty(P, 2, T, dtype=torch.float32, device=dev)
        for p in range(P):
            _lib.check(_lib.load().athtd_chunk_fade_add(seg_out[:, p].data_ptr(), P * 2 * L, 0, L, t_starts.data_ptr(), t_actual.data_ptr(),
                                                        t_fade.data_ptr(), t_flags.data_ptr(), n, stride, up.data_ptr(), down.data_ptr(),
                                                        t_off.data_ptr(), out[p].data_ptr(), 2, 0, T, st), "athtd_chunk_fade_add")
        return out

    # ------------------------------------------------------------------ the reference's plugin interface
    def separate(self, mixture: torch.Tensor, stem_name: str) -> torch.Tensor:
        """benchmark.py:206-208: one stem of a (C, T) mixture."""
        out, _ = self.separate_many(mixture, self.prompt_embeddings([stem_name]))
        return out[0]

    def separate_all(self, mixture: torch.Tensor) -> Dict[str, torch.Tensor]:
        """benchmark.py:210-215: all four stems; here ONE pass (encode once, decode per prompt) instead of four."""
        out, _ = self.separate_many(mixture, self.prompt_embeddings(STEMS))
        return {s: out[i] for i, s in enumerate(STEMS)}
