#!/usr/bin/env python
"""LSH top-k graph on N GPUs (one process per GPU under torchrun): every rank holds all sketches, builds the (replicated)
index and scans all queries, but replays / refines / trims only its own range of neighbour lists (d2g_lsh_topk_rows);
rank 0 gathers the CSR pieces.  No data-path collective: the lists are independent once the candidates are known.
usage: torchrun --nproc-per-node N scripts/topk_multi.py [n] [S] [K]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from dashing2_b200 import capi, synth
from dashing2_b200.shard import equal_rows

n = int(sys.argv[1]) if len(sys.argv) > 1 else 250_000
S = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
K = int(sys.argv[3]) if len(sys.argv) > 3 else 32
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = capi.Context(local)
regs, cards = synth.synthetic_sketches(n, S, seed=5, n_families=max(1, n // 1000))   # same seed on every rank
b = equal_rows(n, world)
x0, x1 = b[rank], b[rank + 1]
ctx.lsh_topk(regs[:2000], cards[:2000], K)                                           # warm up allocations / module load
if world > 1:
    dist.barrier()
torch.cuda.synchronize(); t0 = time.perf_counter()
ip, ix, dv = ctx.lsh_topk(regs, cards, K, rows=(x0, x1))
torch.cuda.synchronize(); dt = time.perf_counter() - t0
stats = torch.tensor([dt, float(ip[-1])], dtype=torch.float64, device="cuda")
if world > 1:
    mx = stats.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    sm = stats.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
else:
    mx = sm = stats
if rank == 0:
    print(f"topk n={n} S={S} K={K} on {world} GPU(s): {float(mx[0])*1e3:.1f} ms (max over ranks, host registers in, CSR rows out)  "
          f"{n/float(mx[0])/1e3:.1f} k sketches/s  nnz={int(sm[1])}", flush=True)
if world > 1:
    dist.destroy_process_group()
