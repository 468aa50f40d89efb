"""bench.py -- audio-seconds separated per second (x realtime) of the segment-batched separation path.

Workload (BASELINE.json configs[2]): a 4-minute synthetic stereo 44.1 kHz track per GPU, 6 s segments with
25 % weighted overlap-add (54 chunks), batch 32, one prompt, bf16 activations / fp32 accumulation, random-init
weights, synthetic unit-norm 512-d text embedding.  One "step" = one complete pass over the track
(gather -> batched forward -> overlap-add).  With N > 1 ranks the track is N x 4 minutes, every rank separates
a contiguous span of chunks and the seams are stitched by one neighbour exchange of the last chunk's output
(weak scaling; no other collective on the data path).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--batch 32] [--seconds 240]

--impl reference times the reference's own algorithm on the host CPU cores (oracle port of
ATHTDemucs_v2.py:250-326 driven by the benchmark.py:155-204 chunk loop, batch 1, all host threads).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch

SR = 44100
GFLOP_PER_SEG_PROMPT = 209.9      # SURVEY.md Appendix C (reference-equivalent, 2 x MAC)
GFLOP_SHARED, GFLOP_PER_PROMPT = 174.1, 35.8


def env_int(name, default):
    return int(os.environ.get(name, default))


class ClockSampler:
    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def cpu_reference_rate(steps: int, warmup: int, chunks_per_step: int, track_seconds: float):
    """x realtime of the oracle port on the host cores, batch 1 per chunk exactly like benchmark.py:155-204."""
    from oracle import athtd_oracle, ola, weights
    torch.set_num_threads(os.cpu_count())
    sd = weights.make_state_dict(0)
    L = 6 * SR
    wav, emb = weights.make_inputs(4, 1, L)
    n_chunks = len(ola.chunk_plan(int(track_seconds * SR)))
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        for _ in range(chunks_per_step):
            athtd_oracle.forward(sd, wav, emb)
        dt = (time.perf_counter() - t0) / chunks_per_step
        if i >= warmup:
            times.append(dt)
    per_chunk = sum(times) / len(times)
    return track_seconds / (n_chunks * per_chunk), per_chunk, n_chunks


def run_reference(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return
    rate, per_chunk, n_chunks = cpu_reference_rate(args.steps, min(args.warmup, 1), 1, args.seconds)
    line = {
        "impl": "reference", "metric": "audio-sec separated/sec (x realtime)", "value": rate, "unit": "x realtime",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": per_chunk * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.seconds:.0f} s synthetic stereo track, 6 s segments, 25% overlap-add ({n_chunks} chunks), 1 prompt",
                   "batch": 1, "note": "oracle port of the reference forward + chunk loop on host CPU; each step = 1 chunk, rate extrapolated to the track"},
        "cpu_baseline": {"value": rate, "unit": "x realtime", "cores": os.cpu_count(), "kind": "port",
                         "sample": f"{args.steps} timed 6 s chunks (batch 1) after 1 warm-up"},
        "e2e": {"value": rate, "unit": "x realtime", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--seconds", type=float, default=240.0)
    ap.add_argument("--prompts", type=int, default=1)
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    # stdout carries exactly one JSON line: everything else that libraries write to file descriptor 1 (NCCL prints its
    # version banner there) goes to stderr; the line is written to the saved descriptor at the end
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    import torch.distributed as dist
    import athtd_b200
    from oracle import weights          # seeded synthetic weights / inputs only (not on the measured path)

    world = env_int("WORLD_SIZE", 1)
    rank = env_int("RANK", 0)
    local = env_int("LOCAL_RANK", 0)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    warmup = max(args.warmup, 3)

    model = athtd_b200.AudioTextHTDemucsB200(precision=args.precision)
    model.load_state_dict(weights.make_state_dict(0), strict=False)
    sep = athtd_b200.B200SeparationModel(model, dev, 6.0, 1.5, batch=args.batch)
    T = int(args.seconds * SR) * world
    g = torch.Generator().manual_seed(4)
    track_host = (0.1 * torch.randn(2, T, generator=g)).pin_memory()
    tt = torch.arange(T) / SR
    track_host += 0.2 * torch.sin(6.2831853 * 220.0 * tt)
    _, emb = weights.make_inputs(2, args.prompts, 4096)
    emb = emb.to(dev)
    plan = athtd_b200.segment_plan(T)
    n = len(plan.starts)
    k0, k1 = athtd_b200.distributed.partition_chunks(n, world)[rank]
    lo, hi = athtd_b200.distributed.span_sample_range(plan.starts, T, (k0, k1))
    in_lo, in_hi = athtd_b200.distributed.span_input_range(plan.starts, T, plan.chunk_len, (k0, k1))
    track_dev = track_host.to(dev)
    out_host = torch.empty(args.prompts, 2, hi - lo).pin_memory()

    halo_exchange = athtd_b200.distributed.make_halo_exchange(rank, world)

    def step(track):
        # span outputs need the left neighbour's last chunk: run own chunks first, exchange, then overlap-add
        return sep.separate_span(track, emb, (k0, k1), halo_exchange)

    def sync():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    def timed(fn, iters):
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        sync()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(warmup):
        step(track_dev)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms = timed(lambda: step(track_dev), args.steps)
    launches = sep.last_launches

    def e2e_step():
        # pinned host track -> batch-wise H2D on a copy stream -> separate -> per-batch overlap-add -> D2H of finished
        # output ranges while the next batch computes (B200SeparationModel.separate_span_host); returns after the last copy
        sep.separate_span_host(track_host, emb, (k0, k1), out_host, halo_exchange)

    e2e_step()
    ms_e2e = timed(e2e_step, args.steps)
    clocks = sampler.stop() if rank == 0 else None

    # dominant-kernel roofline: per-launch CUDA-event timing of the GEMM kernels in a separate (untimed) pass
    prof = sep.profile_gemms(track_dev, emb, (k0, k1))
    audio_s = args.seconds * world
    value = audio_s * args.steps / (ms / 1e3)
    e2e_v = audio_s * args.steps / (ms_e2e / 1e3)
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = peaks.get("bf16_tflops_sustained", 1400.0)
        traffic = None
        try:      # DRAM bytes per launch of the GEMM + attention kernels from the committed ncu capture (B=32 forward)
            tr = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
            traffic = (tr["gemm_tc_kernel"]["dram_bytes_total"] + tr["flash_attn_kernel"]["dram_bytes_total"]) / (
                tr["gemm_tc_kernel"]["launches"] + tr["flash_attn_kernel"]["launches"])
        except Exception:
            pass
        roof = {"bound": "tensor", "kernel": prof["kernel"], "achieved": prof["tflops"], "peak": peak, "unit": "TFLOP/s",
                "frac": prof["tflops"] / peak, "traffic": traffic, "traffic_unit": "bytes per launch (ncu dram read+write, batch-32 forward, profiles/r01_traffic.json)", "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback 1400",
                "launches_per_step": prof["launches"], "gemm_ms_per_step": prof["ms"], "share_of_step": prof["ms"] / (ms / args.steps),
                "algorithmic_gflop_per_step": prof["gflop"]}
        line = {
            "metric": "audio-sec separated/sec (x realtime)", "value": value, "unit": "x realtime", "n_gpus": world,
            "steps": args.steps, "warmup": warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
            "config": {"workload": f"{args.seconds:.0f} s synthetic stereo 44.1 kHz track per GPU, 6 s segments, 25% weighted overlap-add "
                                   f"({n // world} chunks/GPU), batch {args.batch}, {args.prompts} prompt(s), random-init AudioTextHTDemucs",
                       "batch": args.batch, "prompts": args.prompts, "chunks": n, "l2": "inputs+activations per step (>10 GB) exceed the 126 MB L2",
                       "parallelism": f"segment-span x{world}"},
            "stem_seconds_per_s": value * args.prompts,
            "roofline": roof,
            "e2e": {"value": e2e_v, "unit": "x realtime", "h2d_bytes_per_step": int(2 * (in_hi - in_lo) * 4),
                    "d2h_bytes_per_step": int(out_host.numel() * 4), "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches * args.steps,
            "clocks": clocks,
        }
        if not args.no_cpu_baseline:
            rate, per_chunk, nch = cpu_reference_rate(3, 1, 1, args.seconds)
            line["cpu_baseline"] = {"value": rate, "unit": "x realtime", "cores": os.cpu_count(), "kind": "port",
                                    "sample": f"3 timed 6 s chunks (batch 1, {per_chunk * 1e3:.0f} ms each) after 1 warm-up, extrapolated to {nch} chunks"}
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
