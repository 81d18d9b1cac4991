#!/usr/bin/env python
"""Whole-program check on the GPU box: `dashing2 sketch ... --cmpout` (unmodified reference binary, all host threads) vs
`dashing2-gpu` with the same argv on the same FASTA files -- wall clock and byte-level comparison of the stacked sketch
file and the binary distance matrix.  usage: cli_vs_reference.py [n_genomes] [len] [gpus, e.g. 1,2]"""
import os, sys, time, subprocess, shutil, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import refbin
from dashing2_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
L = int(sys.argv[2]) if len(sys.argv) > 2 else 5_000_000
gpu_counts = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [1]
work = tempfile.mkdtemp(prefix="d2cli", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
try:
    paths = synth.write_fasta_set(os.path.join(work, "fa"), n, L, seed=2, n_families=max(1, n // 8))
    flist = os.path.join(work, "files.txt"); open(flist, "w").write("\n".join(paths) + "\n")
    cores = len(os.sched_getaffinity(0))
    gpu = os.path.join(ROOT, "dashing2_b200", "bin", "dashing2-gpu")
    for tag, mode in (("opmh S=1024", ["-S1024"]), ("fss w=51 S=4096", ["-w51", "--full-setsketch", "-S4096"])):
        res = {}
        for who, exe in [("reference", refbin.ref_binary())] + [("gpu" if g == 1 else f"gpu x{g}", gpu) for g in gpu_counts]:
            out = os.path.join(work, who + ".stk"); mat = os.path.join(work, who + ".f32")
            argv = [exe, "sketch", "-k31", "-p", str(cores), "-F", flist, "-o", out, "--binary-output", "--cmpout", mat] + mode
            if who.startswith("gpu x"):
                argv += ["--gpus", who[5:]]
            best = 1e9
            for rep in range(2):
                t0 = time.perf_counter(); r = subprocess.run(argv, capture_output=True, text=True, env=dict(os.environ, OMP_NUM_THREADS=str(cores)))
                best = min(best, time.perf_counter() - t0)
                assert r.returncode == 0, r.stderr[-2000:]
            res[who] = (best, np.fromfile(out, dtype=np.uint8), np.fromfile(mat, dtype=np.float32))
        tr, sr, mr = res["reference"]
        n_hdr = 16 + 8 * n
        for who in res:
            if who == "reference":
                continue
            tg, sg, mg = res[who]
            same_sig = np.array_equal(sr[n_hdr:], sg[n_hdr:]); same_mat = np.array_equal(mr.view(np.uint32), mg.view(np.uint32))
            card_rel = np.max(np.abs(sr[16:n_hdr].view(np.float64) / sg[16:n_hdr].view(np.float64) - 1))
            print(f"{tag}: {n} genomes x {L} bp, k=31: reference {tr:.2f} s ({cores} threads), dashing2-{who} {tg:.2f} s ({tr/tg:.1f}x); "
                  f"registers byte-identical: {same_sig}; matrix byte-identical: {same_mat}; max cardinality rel. diff {card_rel:.1e}", flush=True)
finally:
    shutil.rmtree(work, ignore_errors=True)
