"""bench.py -- audio-seconds separated per second (x realtime) of the segment-batched separation path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config 3|4|5|stft] [--batch 32]

One "step" = one complete pass over the workload's track (gather -> batched forward -> overlap-add).  Workloads are the
BASELINE.json configs:

  --config 3 (default)  a 4-minute synthetic stereo 44.1 kHz track PER GPU, 6 s segments with 25 % weighted overlap-add
                        (54 chunks), batch 32, one prompt, bf16 activations / fp32 accumulation; weak scaling.
  --config 4            256 segments x 4 prompts (T = 256 x 198 450 samples), sharded across the N ranks as contiguous chunk
                        spans with one NVLink halo exchange at each seam; strong scaling (total work fixed).
  --config 5            a 1-hour mix x 8 prompts (800 chunks), same sharding; strong scaling.
  --config stft         STFT -> iSTFT round trip, nfft 4096 / hop 1024, 64 stereo 6 s segments on one GPU (HBM GB/s).

`value` is device-resident throughput (inputs already in HBM); `e2e` is the same pass through the host-staged public call
(`B200SeparationModel.separate_span_host`: pinned host track -> H2D -> separate -> overlap-add -> D2H, copies inside the timed
region, consecutive steps pipelined like a service separating tracks back to back).  With N > 1 ranks every rank separates a
contiguous span of chunks for ALL prompts (encode once, decode per prompt) and the only exchange on the data path is the raw
output of each rank's last chunk to its right neighbour.

--impl reference times the reference's own algorithm on the host CPU cores (the oracle port of ATHTDemucs_v2.py:250-326 driven
by the benchmark.py:155-215 loop: batch 1 per chunk, one full pass per prompt, all host threads), on a bounded sample.
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
STRIDE = 198450                   # 6 s segment - 1.5 s overlap, in samples (benchmark.py:127-128)
GFLOP_PER_SEG_PROMPT = 209.9      # SURVEY.md Appendix C (reference-equivalent, 2 x MAC)
GFLOP_SHARED, GFLOP_PER_PROMPT = 174.1, 35.8
METRIC = "audio-sec separated/sec (x realtime)"


def env_int(name, default):
    return int(os.environ.get(name, default))


def workload(config: str, world: int, seconds_override=None, prompts_override=None):
    """-> (track samples in total, prompts, scaling, description)"""
    if config == "3":
        T, P, scaling = int((seconds_override or 240.0) * SR) * world, 1, "weak"
        what = f"BASELINE config 3: {T / SR / world:.0f} s synthetic stereo 44.1 kHz track per GPU"
    elif config == "4":
        T, P, scaling = 256 * STRIDE, 4, "strong"
        what = "BASELINE config 4: 256 segments x 4 prompts, fixed total work"
    elif config == "5":
        T, P, scaling = 3600 * SR, 8, "strong"
        what = "BASELINE config 5: 1-hour synthetic mix x 8 prompts, fixed total work"
    else:
        raise SystemExit(f"unknown --config {config}")
    if config != "3" and seconds_override:
        T = int(seconds_override * SR)
    if prompts_override:
        P = prompts_override
    return T, P, scaling, what


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


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


# ------------------------------------------------------------------------------------------ reference arm / CPU baseline
def cpu_reference_rate(steps: int, warmup: int, T: int, prompts: int):
    """x realtime of the oracle port on the host cores, batch 1 per chunk and one full pass per prompt exactly like
    benchmark.py:155-215; each timed step is ONE 6 s chunk, the rate is extrapolated to the workload's chunks x prompts."""
    from oracle import athtd_oracle, ola, weights
    torch.set_num_threads(os.cpu_count())
    sd = weights.make_state_dict(0)
    L = 6 * SR
    wav, emb = weights.make_inputs(4, 1, L)
    n_chunks = len(ola.chunk_plan(T))
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        athtd_oracle.forward(sd, wav, emb)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    per_chunk = sum(times) / len(times)
    return (T / SR) / (n_chunks * prompts * per_chunk), per_chunk, n_chunks


def cpu_stft_rate(steps: int, warmup: int, B: int = 4):
    """GB/s (algorithmic bytes) of the oracle's _spec + _ispec on the host cores, B segments per step."""
    from oracle import athtd_oracle, weights
    torch.set_num_threads(os.cpu_count())
    L = 6 * SR
    wav, _ = weights.make_inputs(3, B, L)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        z = athtd_oracle.spec(wav)
        athtd_oracle.ispec(z, L)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    per = sum(times) / len(times)
    Tf = (L + 1023) // 1024
    return 2 * (B * 2 * L * 4 + B * Tf * 2048 * 4 * 4) / per / 1e9, per, B


def run_reference(args):
    if env_int("RANK", 0) != 0:
        return
    world = env_int("WORLD_SIZE", args.gpus)
    warm = min(args.warmup, 1)
    if args.config == "stft":
        gbs, per, B = cpu_stft_rate(args.steps, warm)
        line = {"impl": "reference", "metric": "STFT->iSTFT round trip (algorithmic GB/s)", "value": gbs, "unit": "GB/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": warm, "ms_per_step": per * 1e3, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"BASELINE config 2: STFT->iSTFT round trip, nfft 4096 hop 1024; CPU sample of {B} stereo 6 s segments"},
                "cpu_baseline": {"value": gbs, "unit": "GB/s", "cores": os.cpu_count(), "kind": "port",
                                 "sample": f"{args.steps} timed round trips of {B} segments (torch.stft / istft as the reference calls them)"},
                "e2e": {"value": gbs, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        print(json.dumps(line))
        return
    T, P, scaling, what = workload(args.config, world, args.seconds, args.prompts)
    rate, per_chunk, n_chunks = cpu_reference_rate(args.steps, warm, T, P)
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": "x realtime",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": warm, "ms_per_step": per_chunk * 1e3,
        "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{what}, 6 s segments, 25% weighted overlap-add ({n_chunks} chunks), {P} prompt(s)",
                   "batch": 1, "prompts": P, "chunks": n_chunks,
                   "note": "oracle port of the reference forward + chunk loop on host CPU (batch 1, one full pass per prompt); "
                           "each step = 1 chunk, rate extrapolated to chunks x prompts"},
        "stem_seconds_per_s": rate * P,
        "cpu_baseline": {"value": rate, "unit": "x realtime", "cores": os.cpu_count(), "kind": "port",
                         "sample": f"{args.steps} timed 6 s chunks (batch 1) after {warm} warm-up"},
        "e2e": {"value": rate, "unit": "x realtime", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------ shared GPU-arm plumbing
def pin_rank_cpus(local: int, n_local: int) -> int:
    """Give every local rank its own slice of the host cores (8 launcher processes otherwise migrate over the same cores
    and share their caches / pinned-copy threads).  Returns the number of cores of this rank."""
    try:
        cpus = sorted(os.sched_getaffinity(0))
        per = len(cpus) // max(1, n_local)
        if n_local > 1 and per >= 1:
            os.sched_setaffinity(0, set(cpus[local * per:(local + 1) * per]))
            torch.set_num_threads(per)
            return per
        return len(cpus)
    except Exception:
        return os.cpu_count() or 1


def run_stft(args, json_fd):
    import athtd_b200
    from athtd_b200 import lib as alib
    from athtd_b200 import synthetic
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    B, L = 64, 6 * SR
    Tf = (L + 1023) // 1024
    lib = alib.load()
    eng = athtd_b200.Engine(dev, "fp32")
    wav_host = synthetic.make_inputs(3, B, L)[0].pin_memory()
    out_host = torch.empty(B, 2, L).pin_memory()
    wav = wav_host.to(dev)
    Z = torch.empty(B, Tf, 2048, 4, device=dev)
    frames = torch.empty(B * 2 * Tf * 4096, device=dev)        # scratch argument of the C ABI (unused by the fused inverse)
    out = torch.empty(B, 2, L, device=dev)
    stats = torch.zeros(2 * B, dtype=torch.float64, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    warmup = max(args.warmup, 3)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    acc = {"f": 0.0, "i": 0.0}

    def step(timed=False):
        if timed:
            ev[0].record()
        alib.check(lib.athtd_stft_cac(wav.data_ptr(), B, L, Z.data_ptr(), stats.data_ptr(), eng.tw.data_ptr(), eng.win.data_ptr(), st))
        if timed:
            ev[1].record()
        alib.check(lib.athtd_istft(Z.data_ptr(), B, L, frames.data_ptr(), out.data_ptr(), eng.tw.data_ptr(), eng.win.data_ptr(), st))
        if timed:
            ev[2].record()
            torch.cuda.synchronize()
            acc["f"] += ev[0].elapsed_time(ev[1]); acc["i"] += ev[1].elapsed_time(ev[2])

    def timed(fn, iters):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    for _ in range(warmup):
        step()
    sampler = ClockSampler(0)
    sampler.start()
    ms = timed(step, args.steps) / args.steps
    for _ in range(args.steps):      # per-kernel split with events inside the step (separate pass)
        step(True)

    def e2e_step():
        wav.copy_(wav_host, non_blocking=True)
        step()
        out_host.copy_(out, non_blocking=True)

    e2e_step()
    ms_e2e = timed(e2e_step, args.steps) / args.steps
    clocks = sampler.stop()
    wav_b, z_b = B * 2 * L * 4, B * Tf * 2048 * 4 * 4
    alg = 2 * (wav_b + z_b)
    peaks = measured_peaks()
    peak = peaks.get("hbm_gbs", 6650.0)
    gbs = alg / ms / 1e6
    line = {
        "metric": "STFT->iSTFT round trip (algorithmic GB/s)", "value": gbs, "unit": "GB/s", "n_gpus": 1, "steps": args.steps,
        "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": f"BASELINE config 2: STFT->iSTFT round trip, nfft 4096 hop 1024, {B} stereo 6 s segments, materialised "
                               f"spectrogram [B,Tf,2048,4] fp32", "l2": "every buffer of the round trip (wav 135 MB, Z 543 MB) exceeds the 126 MB L2"},
        "roofline": {"bound": "hbm", "kernel": "stft_cac_kernel + istft kernel(s)", "achieved": gbs, "peak": peak, "unit": "GB/s",
                     "frac": gbs / peak, "traffic": None, "algorithmic_bytes_per_step": alg,
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650",
                     "stft_ms": acc["f"] / args.steps, "istft_ms": acc["i"] / args.steps,
                     "stft_GBs": (wav_b + z_b) / (acc["f"] / args.steps) / 1e6, "istft_GBs": (wav_b + z_b) / (acc["i"] / args.steps) / 1e6,
                     "note": "all bytes counted: read wav + write Z + read Z + write wav; no scratch traffic outside these"},
        "e2e": {"value": alg / ms_e2e / 1e6, "unit": "GB/s", "h2d_bytes_per_step": wav_b, "d2h_bytes_per_step": wav_b, "ms_per_step": ms_e2e},
        "gpu_launches": 3 * args.steps, "clocks": clocks,
    }
    if not args.no_cpu_baseline:
        g, per, nb = cpu_stft_rate(3, 1)
        line["cpu_baseline"] = {"value": g, "unit": "GB/s", "cores": os.cpu_count(), "kind": "port",
                                "sample": f"3 timed round trips of {nb} segments ({per * 1e3:.0f} ms each) after 1 warm-up"}
    os.write(json_fd, (json.dumps(line) + "\n").encode())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--config", default="3", choices=["3", "4", "5", "stft"])
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--seconds", type=float, default=None, help="override the track length (config 3: seconds per GPU)")
    ap.add_argument("--prompts", type=int, default=None, help="override the number of prompts")
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel eagerly (no CUDA-graph replay of the forward)")
    ap.add_argument("--pdl", action="store_true", help="programmatic dependent launch between the kernels (default off: measured slower)")
    ap.add_argument("--no-overlap-batches", action="store_true", help="one batch at a time (default: two batches in flight on two streams / workspaces)")
    ap.add_argument("--no-pin", action="store_true", help="do not give every local rank its own slice of the host cores")
    ap.add_argument("--no-pipeline", action="store_true", help="e2e: wait for the downloads of every step before the next one")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    # stdout carries exactly one JSON line: everything else that libraries write to file descriptor 1 (NCCL prints its
    # version banner there) goes to stderr; the line is written to the saved descriptor at the end
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if args.config == "stft":
        if env_int("RANK", 0) == 0:
            run_stft(args, json_fd)
        return

    import torch.distributed as dist
    import athtd_b200
    from athtd_b200 import synthetic          # seeded synthetic weights / inputs (plain torch on the CPU, outside the timed region)

    world = env_int("WORLD_SIZE", 1)
    rank = env_int("RANK", 0)
    local = env_int("LOCAL_RANK", 0)
    cores = pin_rank_cpus(local, 1 if args.no_pin else env_int("LOCAL_WORLD_SIZE", world))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    warmup = max(args.warmup, 3)

    if args.pdl:
        athtd_b200.load_library().athtd_set_pdl(1)
    model = athtd_b200.AudioTextHTDemucsB200(precision=args.precision)
    model.load_state_dict(synthetic.make_state_dict(0), strict=False)
    sep = athtd_b200.B200SeparationModel(model, dev, 6.0, 1.5, batch=args.batch, use_graph=not args.no_graph,
                                         overlap_batches=not args.no_overlap_batches)
    T, P, scaling, what = workload(args.config, world, args.seconds, args.prompts)
    plan = athtd_b200.segment_plan(T)
    n = len(plan.starts)
    spans = athtd_b200.distributed.partition_chunks(n, world)
    k0, k1 = spans[rank]
    lo, hi = athtd_b200.distributed.span_sample_range(plan.starts, T, (k0, k1))
    in_lo, in_hi = athtd_b200.distributed.span_input_range(plan.starts, T, plan.chunk_len, (k0, k1)) if k1 > k0 else (0, 0)
    # every rank generates only the part of the (seeded) track it reads; the tone keeps global phase
    part_host = synthetic.make_track_part(in_lo, in_hi)
    emb = synthetic.make_prompt_embeddings(P).to(dev)
    track_dev = part_host.to(dev)
    # the host-staged path takes the WHOLE pinned track and reads its span from it: give it a view whose sample 0 is global 0
    out_host = torch.empty(P, 2, hi - lo).pin_memory()
    halo_exchange = athtd_b200.distributed.make_halo_exchange(rank, world, spans)

    def step():
        return sep.separate_span(track_dev, emb, (k0, k1), halo_exchange, track_offset=in_lo, track_len=T)

    def sync():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    def timed(fn, iters, after=None):
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        if after is not None:
            after()
        e1.record()
        sync()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(warmup):
        step()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms = timed(step, args.steps)
    launches = sep.last_launches

    # e2e: pinned host track -> batch-wise H2D on an upload stream -> separate -> per-batch overlap-add -> D2H of finished output
    # ranges on a download stream (B200SeparationModel.separate_span_host).  Consecutive steps are pipelined (the next step's
    # uploads overlap this step's compute and downloads); the timed region ends when the LAST step's downloads have landed.
    host_track = HostTrackView(part_host.pin_memory(), in_lo, T)

    def e2e_step():
        sep.separate_span_host(host_track, emb, (k0, k1), out_host, halo_exchange, wait=args.no_pipeline)

    def e2e_drain():
        if sep.host_done is not None:
            torch.cuda.current_stream(dev).wait_event(sep.host_done)

    for _ in range(warmup):          # untimed: staging buffers, and the graph capture of the (now pointer-stable) forwards
        e2e_step()
    e2e_drain()
    ms_e2e = timed(e2e_step, args.steps, e2e_drain)
    clocks = sampler.stop() if rank == 0 else None

    # dominant-kernel roofline: per-launch CUDA-event timing of the GEMM + attention kernels in a separate (untimed) pass
    prof = sep.profile_gemms(track_dev, emb, (k0, k1), track_offset=in_lo, track_len=T) if k1 > k0 else {"kernel": "", "ms": 0.0, "gflop": 0.0, "launches": 0, "tflops": 0.0}
    audio_s = T / SR
    value = audio_s * args.steps / (ms / 1e3)
    e2e_v = audio_s * args.steps / (ms_e2e / 1e3)
    h2d = torch.tensor([2.0 * (in_hi - in_lo) * 4, float(out_host.numel() * 4)], device=dev)
    if world > 1:
        dist.all_reduce(h2d)
    if rank == 0:
        peaks = measured_peaks()
        peak = peaks.get("bf16_tflops_sustained", 1400.0)
        traffic, traffic_src = None, "not captured for this build (ncu dram bytes are recorded under profiles/ when a capture of the current kernels exists)"
        try:      # DRAM bytes per launch of the GEMM + attention kernels from this round's ncu capture of the same step
            tr = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
            traffic = (tr["gemm_tc_kernel"]["dram_bytes_total"] + tr["flash_attn_kernel"]["dram_bytes_total"]) / (
                tr["gemm_tc_kernel"]["launches"] + tr["flash_attn_kernel"]["launches"])
            traffic_src = "bytes per launch (ncu dram read+write of one batch-32 forward, profiles/r02_traffic.json)"
        except Exception:
            pass
        ms_step = ms / args.steps
        roof = {"bound": "tensor", "kernel": prof["kernel"], "achieved": prof["tflops"], "peak": peak, "unit": "TFLOP/s",
                "frac": prof["tflops"] / peak, "traffic": traffic, "traffic_unit": traffic_src,
                "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback 1400",
                "launches_per_step": prof["launches"], "gemm_ms_per_step": prof["ms"], "share_of_step": prof["ms"] / ms_step if ms_step else None,
                "algorithmic_gflop_per_step": prof["gflop"],
                "reference_equivalent_tflops": GFLOP_PER_SEG_PROMPT * n * P / world / ms_step if ms_step else None,
                "executed_model_tflops": (GFLOP_SHARED + GFLOP_PER_PROMPT * P) * n / world / ms_step if ms_step else None}
        line = {
            "metric": METRIC, "value": value, "unit": "x realtime", "n_gpus": world,
            "steps": args.steps, "warmup": warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": scaling,
            "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
            "config": {"workload": f"{what}, 6 s segments, 25% weighted overlap-add ({n} chunks in total, {k1 - k0} on rank 0), batch {args.batch}, "
                                   f"{P} prompt(s), random-init AudioTextHTDemucs",
                       "batch": args.batch, "prompts": P, "chunks": n, "l2": "inputs+activations per step (>10 GB) exceed the 126 MB L2",
                       "parallelism": f"segment-span x{world}", "host_cores_per_rank": cores,
                       "e2e_pipelined_steps": not args.no_pipeline, "batches_in_flight": 1 if args.no_overlap_batches else 2},
            "stem_seconds_per_s": value * P,
            "roofline": roof,
            "e2e": {"value": e2e_v, "unit": "x realtime", "h2d_bytes_per_step": int(h2d[0].item()),
                    "d2h_bytes_per_step": int(h2d[1].item()), "ms_per_step": ms_e2e / args.steps,
                    "stem_seconds_per_s": e2e_v * P},
            "gpu_launches": launches * args.steps,
            "clocks": clocks,
        }
        if not args.no_cpu_baseline:
            rate, per_chunk, nch = cpu_reference_rate(3, 1, T, P)
            line["cpu_baseline"] = {"value": rate, "unit": "x realtime", "cores": os.cpu_count(), "kind": "port",
                                    "sample": f"3 timed 6 s chunks (batch 1, {per_chunk * 1e3:.0f} ms each) after 1 warm-up, extrapolated to "
                                              f"{nch} chunks x {P} prompt(s)"}
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


class HostTrackView:
    """A rank's pinned part of the track presented as the whole track: ``view[c, a:b]`` with GLOBAL sample indices.  Lets every
    rank pin only the samples it reads (1 hour of stereo fp32 is 1.27 GB) while ``separate_span_host`` keeps its
    whole-track addressing."""

    def __init__(self, part: torch.Tensor, offset: int, T: int):
        self.part, self.offset, self.shape = part, offset, (2, T)

    def __getitem__(self, idx):
        c, sl = idx
        return self.part[c, sl.start - self.offset:sl.stop - self.offset]


if __name__ == "__main__":
    main()
