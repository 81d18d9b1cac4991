#!/usr/bin/env python
"""Golden CSR files for `--nLSH 1` (index of one-register tables only) and `--nLSH 3` (adds 2S four-register tables, three quarters of
them XXH64-keyed; the GPU path has 1 and 2), src/cmp_core.cpp:757-770; UNMODIFIED reference binary, -p1.
Dev container only (needs oracle/_ref).  Same registers as the other top-k goldens (inputs/sk600x64.npz)."""
import os, shutil, sys, tempfile
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__)); ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
from dashing2_b200 import synth  # noqa: E402
import refbin  # noqa: E402
INP = os.path.join(HERE, "inputs"); EXP = os.path.join(HERE, "expected")


def main():
    work = tempfile.mkdtemp(prefix="d2goldn")
    z = np.load(os.path.join(INP, "sk600x64.npz"))
    stk = os.path.join(work, "sk600.ss")
    synth.write_stacked(stk, z["regs"], z["cards"], names=[f"s{i}" for i in range(600)])
    for nlsh in (1, 3, 4, 5):
        for K in (5, 32):
            mat = os.path.join(work, f"top{K}.csr")
            refbin.run_ref(["cmp", "--presketched", "-p1", "--binary-output", "--nLSH", str(nlsh), "--topk", str(K), "--cmpout", mat, stk], threads=1)
            shutil.copy(mat, os.path.join(EXP, f"topk{K}_nlsh{nlsh}_sk600.csr"))
            a = open(mat, "rb").read(); b = open(os.path.join(EXP, f"topk{K}_sk600.csr"), "rb").read()
            print(nlsh, K, len(a), len(b), "differs from nLSH 2:", a != b)
    shutil.rmtree(work)


if __name__ == "__main__":
    main()
