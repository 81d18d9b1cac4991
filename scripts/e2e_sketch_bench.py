#!/usr/bin/env python
"""End-to-end sketch rate of d2g_sketch_batch (pinned host ASCII in, host registers out) for a sweep of upload settings.
usage: e2e_sketch_bench.py [genomes] [len]   env sweeps: D2G_HYBRID_F, D2G_CHUNK_BYTES, D2G_HOST_THREADS"""
import ctypes as C, os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if len(sys.argv) > 3 and sys.argv[3] == "child":
    import numpy as np, torch
    from dashing2_b200 import capi
    import bench
    G, Lg = int(sys.argv[1]), int(sys.argv[2])
    dev = torch.device("cuda", 0)
    ctx = capi.Context(0)
    seq = bench.make_genomes_on_device(torch, dev, G, Lg, seed=2, n_families=max(1, G * 157 // 10000))
    h_seq = torch.empty(G * Lg, dtype=torch.uint8).pin_memory(); h_seq.copy_(seq[:G * Lg]); del seq
    h_off = np.arange(G + 1, dtype=np.uint64) * np.uint64(Lg); h_ent = np.arange(G, dtype=np.uint32)
    S = 4096
    h_sig = torch.empty((G, S), dtype=torch.float64).pin_memory(); h_card = torch.empty(G, dtype=torch.float64).pin_memory()
    p = ctx.params(mode="fss", S=S, k=31, w=51)
    def run():
        nk = C.c_uint64(0)
        rc = ctx.L.d2g_sketch_batch(ctx.h, C.byref(p), h_seq.data_ptr(), h_off.ctypes.data, h_ent.ctypes.data, G, G, None, h_sig.data_ptr(), h_card.data_ptr(), None, C.byref(nk))
        assert rc == 0, ctx.L.d2g_last_error()
        return nk.value
    run()
    ts = []
    for _ in range(6):
        t0 = time.perf_counter(); nk = run(); ts.append(time.perf_counter() - t0)
    print("%-60s %.1f G kmers/s best, %.1f mean (%s s) checksum %.6f" % (os.environ.get("TAG", ""), nk / min(ts) / 1e9, nk * len(ts) / sum(ts) / 1e9,
                                                                      " ".join("%.3f" % t for t in ts), float(h_sig.sum())), flush=True)
else:
    G = sys.argv[1] if len(sys.argv) > 1 else "1024"; Lg = sys.argv[2] if len(sys.argv) > 2 else "5000000"
    sweeps = [{"D2G_HYBRID_F": "1"}, {"D2G_HYBRID_F": "0.7"}, {}, {"D2G_CHUNK_BYTES": str(128 << 20)}]
    if os.environ.get("E2E_FULL_SWEEP"):
        sweeps += [{"D2G_HYBRID_F": "0.85"}, {"D2G_HYBRID_F": "0.5"}, {"D2G_HYBRID_F": "0"}, {"D2G_HYBRID_F": "1", "D2G_HOST_THREADS": "8"}]
    for sw in sweeps:
        env = dict(os.environ, TAG=" ".join(f"{k}={v}" for k, v in sw.items()) or "default (adaptive)", **sw)
        subprocess.run([sys.executable, __file__, G, Lg, "child"], env=env)
