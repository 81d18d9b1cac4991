#!/usr/bin/env python
"""Golden CSR files for `--similarity-threshold x` (NN_GRAPH_THRESHOLD; src/index_build.cpp:53-165 with topk = -1, src/refine.cpp:43-68);
UNMODIFIED reference binary, -p1.  Dev container only (needs oracle/_ref).  Pins the ORACLE; the GPU path is not built yet."""
import os, shutil, sys, tempfile
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__)); ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
from dashing2_b200 import synth  # noqa: E402
import refbin  # noqa: E402
INP = os.path.join(HERE, "inputs"); EXP = os.path.join(HERE, "expected")


def main():
    work = tempfile.mkdtemp(prefix="d2goldt")
    z = np.load(os.path.join(INP, "sk600x64.npz"))
    stk = os.path.join(work, "sk600.ss")
    synth.write_stacked(stk, z["regs"], z["cards"], names=[f"s{i}" for i in range(600)])
    for tag, argv in (("t0.5", ["--similarity-threshold", "0.5"]), ("t0.8", ["--similarity-threshold", "0.8"]),
                      ("t0.3_containment", ["--similarity-threshold", "0.3", "--containment"])):
        mat = os.path.join(work, tag + ".csr")
        refbin.run_ref(["cmp", "--presketched", "-p1", "--binary-output", "--cmpout", mat, stk] + argv, threads=1)
        shutil.copy(mat, os.path.join(EXP, f"nnthr_{tag}_sk600.csr"))
        d = np.fromfile(mat, dtype=np.uint64, count=2)
        print(tag, d)
    shutil.rmtree(work)


if __name__ == "__main__":
    main()
