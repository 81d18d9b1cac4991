#!/usr/bin/env python
"""Timing drivers for the BASELINE configs that are parity-test cases rather than bench lines:
  c3 [genomes] [len] [S] [mode bmh|pmh]  -- counting sketches (--multiset / --prob), device resident
  c4 [n_ref] [n_query] [S]               -- panel compare, device resident (rows = refs, cols = queries)
  c5 [n] [S] [K]                         -- LSH top-k graph, host registers in, CSR out
Sketch matrices are synthesised on the device (family base row + per-register resampling)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dashing2_b200 import capi
import bench
cmd = sys.argv[1]
args = [int(x) if x.isdigit() else x for x in sys.argv[2:]]
ctx = capi.Context(0); dev = torch.device("cuda", 0); ext = torch.cuda.ExternalStream(ctx.stream)


def sketches_on_device(n, S, seed, n_fam):
    g = torch.Generator(device=dev); g.manual_seed(seed)
    base = torch.rand((n_fam, S), dtype=torch.float64, device=dev, generator=g)
    out = torch.empty((n, S), dtype=torch.float64, device=dev)
    step = 8192
    for i in range(0, n, step):
        m = min(step, n - i)
        fam = (torch.arange(i, i + m, device=dev) % n_fam)
        p = 0.05 + 0.9 * torch.rand((m, 1), dtype=torch.float64, device=dev, generator=g)
        fresh = torch.rand((m, S), dtype=torch.float64, device=dev, generator=g)
        keep = torch.rand((m, S), dtype=torch.float64, device=dev, generator=g) >= p
        out[i:i + m] = torch.where(keep, base[fam], fresh)
    return out, torch.full((n,), 1e6, dtype=torch.float64, device=dev)


def timed(fn, reps=2):
    best = None
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(ext); fn(); e1.record(ext); ext.synchronize()
        ms = e0.elapsed_time(e1); best = ms if best is None else min(best, ms)
    return best


if cmd == "c3":
    G, Lg, S, mode = (args + [8, 20_000_000, 8192, "bmh"][len(args):])[:4]
    seq = bench.make_genomes_on_device(torch, dev, G, Lg, seed=3, n_families=max(1, G // 16))
    # 10 % of each genome duplicated once so that counts > 1 exist: append a copy of the first tenth as a second record
    rec_off = torch.arange(G + 1, dtype=torch.int64, device=dev) * Lg
    rec_ent = torch.arange(G, dtype=torch.int32, device=dev)
    sig = torch.empty((G, S), dtype=torch.float64, device=dev); card = torch.empty(G, dtype=torch.float64, device=dev)
    p = ctx.params(mode=mode, S=S, k=31)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for rep in range(2):
        t0 = time.perf_counter()
        ctx.sketch_batch_dev(p, seq.data_ptr(), rec_off.data_ptr(), rec_ent.data_ptr(), G, G, G * Lg, sig_d=sig.data_ptr(), card_d=card.data_ptr())
        ctx.sync(); dt = time.perf_counter() - t0
        print(f"c3 {mode} G={G} L={Lg} S={S}: {dt*1e3:.1f} ms  {G*(Lg-30)/dt/1e9:.3f} G kmers/s  card[0]={float(card[0]):.1f} sigsum={float(sig.sum()):.6g}", flush=True)
elif cmd == "c4":
    nf, nq, S = (args + [50_000, 100_000, 1024][len(args):])[:3]
    regs, cards = sketches_on_device(nf + nq, S, 4, max(1, (nf + nq) // 100))
    p = ctx.cmp_params(S, nf + nq, "panel", "similarity", k=31, nq=nq)
    out = torch.empty(nf * nq, dtype=torch.float32, device=dev)
    ctx.set_timing(True)
    ms = timed(lambda: ctx.cmp_rows_dev(p, regs.data_ptr(), cards.data_ptr(), 0, nf, out.data_ptr()))
    tile_ms, tile_n = ctx.get_timing(2); prep_ms, _ = ctx.get_timing(3)
    print(f"c4 panel {nf} x {nq} S={S}: {ms:.1f} ms  {nf*nq/ms/1e6:.2f} G pairs/s  (2 reps: tile kernels {tile_ms:.1f} ms in {tile_n} launches, code prep {prep_ms:.1f} ms)  "
          f"mean sim={float(out[::9973].double().mean()):.5f}", flush=True)
elif cmd == "c5":
    n, S, K = (args + [250_000, 1024, 32][len(args):])[:3]
    regs, cards = sketches_on_device(n, S, 5, max(1, n // 1000))
    h_regs_t = torch.empty(regs.shape, dtype=torch.float64).pin_memory(); h_regs_t.copy_(regs)     # page-locked host registers
    h_regs = h_regs_t.numpy(); h_cards = cards.cpu().numpy()
    del regs; torch.cuda.empty_cache()
    for rep in range(2):
        l0 = ctx.launch_count(); t0 = time.perf_counter()
        ip, ix, dv = ctx.lsh_topk(h_regs, h_cards, K)
        dt = time.perf_counter() - t0
        print(f"c5 topk n={n} S={S} K={K}: {dt*1e3:.1f} ms wall (host registers in, CSR out)  {n/dt/1e3:.1f} k sketches/s  nnz={int(ip[-1])}  launches={ctx.launch_count()-l0}", flush=True)
