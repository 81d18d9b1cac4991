#!/usr/bin/env python
"""Phase timing of the front-end (`dashing2-gpu -v`) on FASTA files in /dev/shm.  usage: cli_phases.py [n_genomes] [len]"""
import os, sys, time, subprocess, shutil, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dashing2_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
L = int(sys.argv[2]) if len(sys.argv) > 2 else 5_000_000
work = tempfile.mkdtemp(prefix="d2cli", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
try:
    paths = synth.write_fasta_set(os.path.join(work, "fa"), n, L, seed=2, n_families=max(1, n // 8))
    flist = os.path.join(work, "files.txt"); open(flist, "w").write("\n".join(paths) + "\n")
    cores = len(os.sched_getaffinity(0))
    gpu = os.path.join(ROOT, "dashing2_b200", "bin", "dashing2-gpu")
    for tag, mode in (("opmh S=1024", ["-S1024"]), ("fss w=51 S=4096", ["-w51", "--full-setsketch", "-S4096"])):
        for rep in range(3):
            out = os.path.join(work, "o.stk"); mat = os.path.join(work, "o.f32")
            argv = [gpu, "sketch", "-v", "-k31", "-p", str(cores), "-F", flist, "-o", out, "--binary-output", "--cmpout", mat] + mode
            t0 = time.perf_counter(); r = subprocess.run(argv, capture_output=True, text=True); dt = time.perf_counter() - t0
            print(f"== {tag} rep {rep}: {dt:.2f} s wall ({cores} threads)\n{r.stderr}", flush=True)
finally:
    shutil.rmtree(work, ignore_errors=True)
