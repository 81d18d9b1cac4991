#!/usr/bin/env python
"""Golden vectors for k > 32 (RollingHasher / CyclicHash k-mer hashes, bonsai encoder.h:644-865); UNMODIFIED reference binary.
Dev container only (needs oracle/_ref).  Pins the ORACLE's rolling-hash stream; the GPU path for k > 32 is not built yet."""
import gzip, os, shutil, sys, tempfile
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__)); ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import refbin  # noqa: E402
from make_golden import read_stacked  # noqa: E402
INP = os.path.join(HERE, "inputs"); EXP = os.path.join(HERE, "expected")
CASES = {
    "roll_opmh_k40_S128":        ["-k40", "-S128"],
    "roll_opmh_k64_S64_nocanon": ["-k64", "-S64", "-C"],
    "roll_opmh_k40_w60_S64":     ["-k40", "-w60", "-S64"],
    "roll_opmh_k33_w50_S64_nocanon": ["-k33", "-w50", "-S64", "-C"],
    "roll_fss_k45_S64_seed3":    ["-k45", "-S64", "--full-setsketch", "--seed", "3"],
}
FILES = ["g0.fa", "g1.fa", "dup.fa", "adv.fa", "reads.fq"]


def main():
    work = tempfile.mkdtemp(prefix="d2goldr")
    paths = []
    for n in FILES:
        dst = os.path.join(work, n); open(dst, "wb").write(gzip.open(os.path.join(INP, n + ".gz"), "rb").read()); paths.append(dst)
    for name, argv in CASES.items():
        out = os.path.join(work, name + ".stk")
        refbin.run_ref(["sketch", "-p1", "-o", out] + argv + paths, threads=1)
        cards, sigs = read_stacked(out)
        np.savez_compressed(os.path.join(EXP, name + ".npz"), cards=cards, sigs=sigs)
        print(name, cards)
    shutil.rmtree(work)


if __name__ == "__main__":
    main()
