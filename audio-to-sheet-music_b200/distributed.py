"""Multi-GPU sharding of the track-level path: one process per GPU (torch.distributed, NCCL over NVLink on the
B200 box, gloo in the CPU tests), contiguous chunk spans per rank, ONE neighbour exchange at the seams.

The reference has no distributed code; this is the B200-native scaling axis of
`OurModel._chunked_inference` (/root/reference/benchmark.py:155-204): every chunk is independent and every
output sample has at most two contributing chunks, so rank r only needs the raw model output of rank r-1's last
chunk to overlap-add its own span bit-identically to the serial loop.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import torch
import torch.distributed as dist


def partition_chunks(n_chunks: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous, balanced chunk spans [(k0, k1)] per rank (rank r: [r*n//R, (r+1)*n//R))."""
    return [(r * n_chunks // world, (r + 1) * n_chunks // world) for r in range(world)]


def span_sample_range(starts: List[int], T: int, span: Tuple[int, int]) -> Tuple[int, int]:
    """Track samples a rank owns after the stitch: [starts[k0], starts[k1]) (to T for the last span)."""
    k0, k1 = span
    return starts[k0], (T if k1 == len(starts) else starts[k1])


def span_input_range(starts: List[int], T: int, chunk_len: int, span: Tuple[int, int]) -> Tuple[int, int]:
    """Track samples a rank must read to run its chunks."""
    k0, k1 = span
    return starts[k0], min(T, starts[k1 - 1] + chunk_len)


def make_halo_exchange(rank: int, world: int, spans: Optional[List[Tuple[int, int]]] = None,
                       group=None) -> Callable[[torch.Tensor], Optional[torch.Tensor]]:
    """Returns f(last_chunk_out) -> left neighbour's last chunk output (None on rank 0 / single rank).
    Ranks with an empty span forward nothing; `spans` lets them be skipped."""

    def nonempty(r):
        return spans is None or spans[r][1] > spans[r][0]

    def exchange(last_out: torch.Tensor) -> Optional[torch.Tensor]:
        if world == 1:
            return None
        right = next((r for r in range(rank + 1, world) if nonempty(r)), None)
        left = next((r for r in range(rank - 1, -1, -1) if nonempty(r)), None)
        ops, recv = [], None
        if right is not None and nonempty(rank):
            ops.append(dist.P2POp(dist.isend, last_out.contiguous(), right, group=group))
        if left is not None and nonempty(rank):
            recv = torch.empty_like(last_out)
            ops.append(dist.P2POp(dist.irecv, recv, left, group=group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        return recv

    return exchange
