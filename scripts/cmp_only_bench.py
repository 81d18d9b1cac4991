#!/usr/bin/env python
"""Small driver used under ncu: runs only the compare kernel on n synthetic sketches (device resident)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dashing2_b200 import capi, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
S = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
ctx = capi.Context(0)
regs, cards = synth.synthetic_sketches(n, S, seed=4, n_families=max(1, n // 64))
r = torch.from_numpy(regs).cuda(); c = torch.from_numpy(cards).cuda()
p = ctx.cmp_params(S, n, "symmetric", "similarity", k=31)
out = torch.empty(n * (n - 1) // 2, dtype=torch.float32, device="cuda")
torch.cuda.synchronize()
ext = torch.cuda.ExternalStream(ctx.stream)
for i in range(reps):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(ext); ctx.cmp_rows_dev(p, r.data_ptr(), c.data_ptr(), 0, n, out.data_ptr()); e1.record(ext); ext.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"n={n} S={S} {ms:.3f} ms  {n*(n-1)/2/ms/1e3:.1f} Mpairs/s  {n*(n-1)/2*S/ms/1e9:.2f} Treg-cmp/s")
