#!/usr/bin/env python
"""Small driver (also used under ncu): Full SetSketch / OPMH sketching of G synthetic genomes, device resident.
usage: sketch_only_bench.py [genomes] [len] [reps] [mode: fss|opmh|pmh|bmh] [S] [w]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
G = int(sys.argv[1]) if len(sys.argv) > 1 else 64
Lg = int(sys.argv[2]) if len(sys.argv) > 2 else 5_000_000
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
mode = sys.argv[4] if len(sys.argv) > 4 else "fss"
S = int(sys.argv[5]) if len(sys.argv) > 5 else 4096
w = int(sys.argv[6]) if len(sys.argv) > 6 else 51
import torch
from dashing2_b200 import capi
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench
ctx = capi.Context(0)
dev = torch.device("cuda", 0)
seq = bench.make_genomes_on_device(torch, dev, G, Lg, seed=2, n_families=max(1, G * 157 // 10000))
rec_off = torch.arange(G + 1, dtype=torch.int64, device=dev) * Lg
rec_ent = torch.arange(G, dtype=torch.int32, device=dev)
sig = torch.empty((G, S), dtype=torch.float64, device=dev); card = torch.empty(G, dtype=torch.float64, device=dev)
m = S + (S & 1)
regs = torch.empty((G, m), dtype=torch.int64, device=dev)
p = ctx.params(mode=mode, S=S, k=31, w=w)
ext = torch.cuda.ExternalStream(ctx.stream)
torch.cuda.synchronize()
ctx.set_timing(True)
for i in range(reps):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(ext)
    if mode == "opmh":
        ctx.sketch_batch_dev(p, seq.data_ptr(), rec_off.data_ptr(), rec_ent.data_ptr(), G, G, G * Lg, regs_u64_d=regs.data_ptr())
    else:
        ctx.sketch_batch_dev(p, seq.data_ptr(), rec_off.data_ptr(), rec_ent.data_ptr(), G, G, G * Lg, sig_d=sig.data_ptr(), card_d=card.data_ptr())
    e1.record(ext); ext.synchronize()
    ms = e0.elapsed_time(e1)
    main_ms, _ = ctx.get_timing(0); boot_ms, _ = ctx.get_timing(1)
    chk = float(sig.sum()) if mode != "opmh" else int(regs.sum())
    print(f"mode={mode} G={G} L={Lg} S={S} w={w}: {ms:.3f} ms total (main {main_ms:.3f} ms, boot/other {boot_ms:.3f} ms)  {G*(Lg-30)/ms/1e6:.2f} G kmers/s  checksum={chk}")
