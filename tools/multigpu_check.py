"""torchrun --nproc-per-node N tools/multigpu_check.py : N-rank span separation (NCCL halo exchange) vs the single-rank
result computed on rank 0, must be bit-identical (SURVEY.md 8e / Appendix H last row)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import athtd_b200
from athtd_b200 import distributed as adist
from oracle import weights

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 4      # 2: several batches per rank (two in flight) next to the halo exchange
model = athtd_b200.AudioTextHTDemucsB200(precision=prec)
model.load_state_dict(weights.make_state_dict(0), strict=False)
sep = athtd_b200.B200SeparationModel(model, dev, 6.0, 1.5, batch=batch)
T = 198450 * 3 * world + 12345                       # a few chunks per rank, ragged tail
wav, _ = weights.make_inputs(77, 1, T)
mix = wav[0].to(dev)
emb = torch.stack([weights.make_inputs(80 + p, 1, 4096)[1][0] for p in range(2)]).to(dev)      # 2 prompts
plan = athtd_b200.segment_plan(T)
spans = adist.partition_chunks(len(plan.starts), world)
halo = adist.make_halo_exchange(rank, world, spans)
in_lo, in_hi = adist.span_input_range(plan.starts, T, plan.chunk_len, spans[rank])
part = sep.separate_span(mix[:, in_lo:in_hi].contiguous(), emb, spans[rank], halo, track_offset=in_lo, track_len=T)
lo, hi = adist.span_sample_range(plan.starts, T, spans[rank])
full = torch.zeros(2, 2, T, device=dev)
full[:, :, lo:hi] = part
dist.all_reduce(full)                                   # disjoint spans: sum == concatenation
ok = True
if rank == 0:
    ref, _ = sep.separate_many(mix, emb)
    ok = bool(torch.equal(ref, full))
    print(f"multigpu_check world={world} precision={prec} batch={batch} chunks={len(plan.starts)} bit_exact={ok} max|diff|={(ref - full).abs().max().item():.3e}")
flag = torch.tensor([1 if ok else 0], device=dev)
dist.broadcast(flag, 0)
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) == 1 else 1)
