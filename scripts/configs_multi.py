#!/usr/bin/env python
"""BASELINE configs 4 and 5 on N GPUs of one box (one process per GPU under torchrun; MAX over ranks of device / wall time).

  c4 [n_ref] [n_query] [S]  panel compare: every rank synthesises 1/N of the sketches (refs first, then queries), ONE NCCL all-gather
                            of the register matrix, then each rank computes an equal range of the |F| rows of the |F| x |Q| float32
                            matrix (d2g_cmp_rows_dev, shape PANEL).  No other collective.
  c5 [n] [S] [K]            LSH top-k graph: every rank holds all sketches (page-locked host registers in), builds the replicated index,
                            scans all queries and replays / refines / trims its own range of lists (d2g_lsh_topk_rows).  No collective.
usage: torchrun --nproc-per-node N scripts/configs_multi.py c4|c5 [...]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from dashing2_b200 import capi
from dashing2_b200.shard import equal_rows

cmd = sys.argv[1]
args = [int(x) for x in sys.argv[2:]]
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
ctx = capi.Context(local); ext = torch.cuda.ExternalStream(ctx.stream, device=dev)


def sketches_on_device(i0, i1, S, seed, n_fam):
    """rows [i0, i1) of the synthetic register matrix of SURVEY 8(d): family base row + per-register resampling."""
    g = torch.Generator(device=dev); g.manual_seed(seed)
    base = torch.rand((n_fam, S), dtype=torch.float64, device=dev, generator=g)
    g.manual_seed(seed * 1000003 + i0)
    out = torch.empty((i1 - i0, S), dtype=torch.float64, device=dev)
    for i in range(i0, i1, 8192):
        m = min(8192, i1 - i)
        fam = torch.arange(i, i + m, device=dev) % n_fam
        p = 0.05 + 0.9 * torch.rand((m, 1), dtype=torch.float64, device=dev, generator=g)
        fresh = torch.rand((m, S), dtype=torch.float64, device=dev, generator=g)
        keep = torch.rand((m, S), dtype=torch.float64, device=dev, generator=g) >= p
        out[i - i0:i - i0 + m] = torch.where(keep, base[fam], fresh)
    return out


def maxr(x):
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


if cmd == "c4":
    nf, nq, S = (args + [50_000, 100_000, 1024][len(args):])[:3]
    n = nf + nq
    sb = equal_rows(n, world)
    if any(sb[r + 1] - sb[r] != sb[1] - sb[0] for r in range(world)):
        raise SystemExit("c4: n must divide by the number of ranks (all_gather_into_tensor)")
    mine = sketches_on_device(sb[rank], sb[rank + 1], S, 4, max(1, n // 100))
    regs = torch.empty((n, S), dtype=torch.float64, device=dev)
    cards = torch.full((n,), 1e6, dtype=torch.float64, device=dev)
    rb = equal_rows(nf, world); r0, r1 = rb[rank], rb[rank + 1]
    p = ctx.cmp_params(S, n, "panel", "similarity", k=31, nq=nq)
    out = torch.empty(ctx.cmp_rows_size(p, r0, r1), dtype=torch.float32, device=dev)
    res = []
    for rep in range(3):
        barrier()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        e[0].record()
        if world > 1:
            dist.all_gather_into_tensor(regs, mine)
        else:
            regs.copy_(mine)
        e[1].record(); torch.cuda.current_stream().synchronize()
        e[2].record(ext)
        ctx.cmp_rows_dev(p, regs.data_ptr(), cards.data_ptr(), r0, r1, out.data_ptr())
        e[3].record(ext); ext.synchronize()
        res.append((maxr(e[0].elapsed_time(e[1])), maxr(e[2].elapsed_time(e[3]))))
    ag, cm = min(res[1:], key=lambda t: t[0] + t[1])
    chk = maxr(float(out[::9973].double().mean()))
    if rank == 0:
        print(f"c4 panel {nf} x {nq} S={S} on {world} GPU(s): all-gather {ag:.1f} ms + compare {cm:.1f} ms (max over ranks, device resident)  "
              f"{nf*nq/(ag+cm)/1e6:.2f} G pairs/s  ({nf*nq/cm/1e6:.2f} G pairs/s compare only)  mean sim sample={chk:.5f}", flush=True)
elif cmd == "c5":
    n, S, K = (args + [1_000_000, 1024, 32][len(args):])[:3]
    h_regs_t = torch.empty((n, S), dtype=torch.float64).pin_memory()
    step = 1 << 17
    for i in range(0, n, step):                     # same seeds on every rank -> the same matrix on every rank
        j = min(n, i + step)
        h_regs_t[i:j].copy_(sketches_on_device(i, j, S, 5, max(1, n // 1000)))
    torch.cuda.empty_cache()
    h_regs = h_regs_t.numpy(); h_cards = np.full(n, 1e6)
    b = equal_rows(n, world); x0, x1 = b[rank], b[rank + 1]
    ctx.lsh_topk(h_regs[:2000], h_cards[:2000], K)  # warm up allocations / module load
    best = None
    for rep in range(2):
        barrier(); t0 = time.perf_counter()
        ip, ix, dv = ctx.lsh_topk(h_regs, h_cards, K, rows=(x0, x1))
        torch.cuda.synchronize(); dt = maxr(time.perf_counter() - t0)
        best = dt if best is None else min(best, dt)
    nnz = torch.tensor([float(ip[-1])], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(nnz)
    if rank == 0:
        print(f"c5 topk n={n} S={S} K={K} on {world} GPU(s): {best*1e3:.1f} ms (max over ranks, page-locked host registers in, CSR rows out)  "
              f"{n/best/1e3:.1f} k sketches/s  nnz={int(nnz[0])}", flush=True)
else:
    raise SystemExit(__doc__)
if world > 1:
    dist.destroy_process_group()
