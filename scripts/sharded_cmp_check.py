#!/usr/bin/env python
"""Under torchrun (one rank per GPU): d2g_cmp_rows_sharded_dev over the library's own NCCL communicator against the single-GPU
d2g_cmp_rows_dev on all-gathered registers, bit for bit, plus timing of both.  usage: torchrun ... sharded_cmp_check.py [n_per_rank] [S]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from dashing2_b200 import capi, synth
from dashing2_b200.shard import equal_area_rows

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n_per = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
S = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
n = n_per * world
ctx = capi.Context(local)
uid = [ctx.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(uid, src=0)
ctx.comm_init_rank(world, rank, uid[0])
regs, cards = synth.synthetic_sketches(n_per, S, seed=100 + rank, n_families=max(2, n_per // 64))
t_regs = torch.from_numpy(regs).to(dev); t_cards = torch.from_numpy(cards).to(dev)
all_regs = torch.empty((n, S), dtype=torch.float64, device=dev); all_cards = torch.empty(n, dtype=torch.float64, device=dev)
dist.all_gather_into_tensor(all_regs, t_regs); dist.all_gather_into_tensor(all_cards, t_cards)
ok = True
for shape, measure in (("symmetric", "similarity"), ("asymmetric", "containment")):
    p = ctx.cmp_params(S, n, shape, measure, k=31)
    b = equal_area_rows(n, world) if shape == "symmetric" else [n * i // world for i in range(world + 1)]
    r0, r1 = b[rank], b[rank + 1]
    nv = ctx.cmp_rows_size(p, r0, r1)
    out_a = torch.empty(nv, dtype=torch.float32, device=dev); out_b = torch.empty(nv, dtype=torch.float32, device=dev)
    ext = torch.cuda.ExternalStream(ctx.stream, device=dev)
    for it in range(3):
        torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
        ctx.cmp_rows_sharded_dev(p, t_regs.data_ptr(), t_cards.data_ptr(), rank * n_per, n_per, r0, r1, out_a.data_ptr())
        ctx.sync(); ta = time.perf_counter() - t0
        torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
        ctx.cmp_rows_dev(p, all_regs.data_ptr(), all_cards.data_ptr(), r0, r1, out_b.data_ptr())
        ctx.sync(); tb = time.perf_counter() - t0
    same = bool(torch.equal(out_a.view(torch.int32), out_b.view(torch.int32)))
    ok &= same
    t = torch.tensor([ta, tb], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"{shape}/{measure}: n={n} S={S} ranks={world}: sharded {t[0].item()*1e3:.2f} ms (exchange + 1/{world} of the ranking + rows), "
              f"replicated prep {t[1].item()*1e3:.2f} ms (registers already gathered); identical={same}", flush=True)
flag = torch.tensor([int(ok)], device=dev); dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("SHARDED_CMP_OK" if flag.item() else "SHARDED_CMP_MISMATCH", flush=True)
dist.destroy_process_group()
