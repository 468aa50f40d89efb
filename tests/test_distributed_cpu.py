"""CPU, world_size 2 (gloo): the multi-rank host logic -- span partition, one neighbour halo exchange, per-span
overlap-add -- reproduces the single-process reference loop bit for bit.  The compute here is the ORACLE's stand-in
model and OLA (checker only); the product pieces under test are athtd_b200.distributed and segment_plan."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import athtd_b200
from athtd_b200 import distributed as adist
from oracle import ola


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _standin(chunk):
    return torch.tanh(chunk * 3.0) + 0.25 * chunk


def _span_ola(plan, seg_outs, k_base, t_begin, t_end):
    """per-span overlap-add in the reference's operation order (oracle-side checker)."""
    out = torch.zeros(2, t_end - t_begin)
    wsum = torch.zeros(t_end - t_begin)
    chunks = ola.chunk_plan(plan.T, 6.0, 1.5)
    for k, c in enumerate(chunks):
        if k < k_base or k - k_base >= len(seg_outs) or seg_outs[k - k_base] is None:
            continue
        w = ola.chunk_weight(c)
        lo, hi = max(c.start, t_begin), min(c.end, t_end)
        if lo >= hi:
            continue
        o = seg_outs[k - k_base][:, lo - c.start:hi - c.start]
        out[:, lo - t_begin:hi - t_begin] += o * w[lo - c.start:hi - c.start]
        wsum[lo - t_begin:hi - t_begin] += w[lo - c.start:hi - c.start]
    return out / wsum.clamp(min=1e-8)


def _worker(rank, world, port, T, result_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(T)
    mix = torch.randn(2, T, generator=g)
    plan = athtd_b200.segment_plan(T)
    spans = adist.partition_chunks(len(plan.starts), world)
    k0, k1 = spans[rank]
    in_lo, in_hi = adist.span_input_range(plan.starts, T, plan.chunk_len, (k0, k1))
    part = mix[:, in_lo:in_hi]
    outs = []
    for k in range(k0, k1):
        seg = torch.zeros(2, plan.chunk_len)
        n = min(plan.chunk_len, T - plan.starts[k])
        seg[:, :n] = part[:, plan.starts[k] - in_lo:plan.starts[k] - in_lo + n]
        outs.append(_standin(seg)[:, :plan.actual_len[k]])
    halo_fn = adist.make_halo_exchange(rank, world, spans)
    last = torch.zeros(2, plan.chunk_len); last[:, :outs[-1].shape[1]] = outs[-1]
    halo = halo_fn(last)
    t_begin, t_end = adist.span_sample_range(plan.starts, T, (k0, k1))
    seg_outs = ([halo[:, :plan.actual_len[k0 - 1]]] if k0 > 0 else [None]) + outs
    res = _span_ola(plan, seg_outs, k0 - 1, t_begin, t_end)
    torch.save((t_begin, t_end, res), os.path.join(result_dir, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("T", [600000, 1000003])
def test_two_rank_span_stitch_is_bit_exact(tmp_path, T):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), T, str(tmp_path)), nprocs=world, join=True)
    parts = [torch.load(os.path.join(tmp_path, f"r{r}.pt")) for r in range(world)]
    assert parts[0][0] == 0 and parts[0][1] == parts[1][0] and parts[1][1] == T
    got = torch.cat([p[2] for p in parts], dim=1)
    g = torch.Generator().manual_seed(T)
    mix = torch.randn(2, T, generator=g)
    ref = ola.chunked_inference(_standin, mix)
    assert torch.equal(got, ref)


def test_partition_and_ranges():
    plan = athtd_b200.segment_plan(50803200)            # BASELINE config 4: exactly 256 chunks
    assert len(plan.starts) == 256
    for world in (1, 2, 4, 8):
        spans = adist.partition_chunks(256, world)
        assert spans[0][0] == 0 and spans[-1][1] == 256
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert all(k1 - k0 == 256 // world for k0, k1 in spans)
        cover = [adist.span_sample_range(plan.starts, plan.T, s) for s in spans]
        assert cover[0][0] == 0 and cover[-1][1] == plan.T and all(a[1] == b[0] for a, b in zip(cover, cover[1:]))
