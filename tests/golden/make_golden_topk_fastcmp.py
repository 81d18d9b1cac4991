#!/usr/bin/env python
"""Golden CSR files for `--topk K` together with `--fastcmp N [--bbit-sigs]` (index over the f64 signatures, refinement through the
compressed compare branch); UNMODIFIED reference binary, -p1.  Dev container only.  Pins the ORACLE; libd2gpu rejects the combination."""
import os, shutil, sys, tempfile
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__)); ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
from dashing2_b200 import synth  # noqa: E402
import refbin  # noqa: E402
INP = os.path.join(HERE, "inputs"); EXP = os.path.join(HERE, "expected")


def main():
    work = tempfile.mkdtemp(prefix="d2goldtf")
    z = np.load(os.path.join(INP, "sk600x64.npz"))
    stk = os.path.join(work, "sk600.ss")
    synth.write_stacked(stk, z["regs"], z["cards"], names=[f"s{i}" for i in range(600)])
    for tag, argv in (("fd1", ["--fastcmp", "1"]), ("fd2_bbit", ["--fastcmp", "2", "--bbit-sigs"])):
        mat = os.path.join(work, tag + ".csr")
        refbin.run_ref(["cmp", "--presketched", "-p1", "--binary-output", "--topk", "8", "--cmpout", mat, stk] + argv, threads=1)
        shutil.copy(mat, os.path.join(EXP, f"topk8_{tag}_sk600.csr"))
        print(tag, np.fromfile(mat, dtype=np.uint64, count=2))
    shutil.rmtree(work)


if __name__ == "__main__":
    main()
