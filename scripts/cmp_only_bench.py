#!/usr/bin/env python
"""Small driver (also used under ncu): runs only the compare path on n synthetic sketches, device resident.
usage: cmp_only_bench.py [n] [S] [reps] [path: codes|f64|auto] [measure]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
S = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
path = sys.argv[4] if len(sys.argv) > 4 else "auto"
measure = sys.argv[5] if len(sys.argv) > 5 else "similarity"
if path != "auto":
    os.environ["D2G_CMP_PATH"] = path
import torch
from dashing2_b200 import capi, synth
ctx = capi.Context(0)
regs, cards = synth.synthetic_sketches(n, S, seed=4, n_families=max(1, n // 64))
r = torch.from_numpy(regs).cuda(); c = torch.from_numpy(cards).cuda()
p = ctx.cmp_params(S, n, "symmetric", measure, k=31)
out = torch.empty(n * (n - 1) // 2, dtype=torch.float32, device="cuda")
torch.cuda.synchronize()
ext = torch.cuda.ExternalStream(ctx.stream)
ctx.set_timing(True)
for i in range(reps):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(ext); ctx.cmp_rows_dev(p, r.data_ptr(), c.data_ptr(), 0, n, out.data_ptr()); e1.record(ext); ext.synchronize()
    ms = e0.elapsed_time(e1)
    tile_ms, tile_n = ctx.get_timing(2); prep_ms, prep_n = ctx.get_timing(3)
    print(f"path={path} n={n} S={S} {measure}: {ms:.3f} ms total  (tile kernel {tile_ms:.3f} ms x{tile_n}, code prep {prep_ms:.3f} ms)  "
          f"{n*(n-1)/2/ms/1e3:.1f} Mpairs/s  {n*(n-1)/2*S/ms/1e9:.2f} Treg-cmp/s  checksum={float(out.double().sum()):.6f}")
