#!/usr/bin/env python
"""Golden vectors for the counting sketches over a count sketch (`--countsketch-size n`, Counter::add / finalize, src/counter.h:68-77,131-137);
UNMODIFIED reference binary.  Dev container only (needs oracle/_ref)."""
import gzip, os, shutil, sys, tempfile
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__)); ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import refbin  # noqa: E402
from make_golden import read_stacked  # noqa: E402
INP = os.path.join(HERE, "inputs"); EXP = os.path.join(HERE, "expected")
CASES = {
    "cs5000_bmh_k31_S32":   ["-k31", "-S32", "--multiset", "--countsketch-size", "5000"],
    "cs5000_pmh_k31_S32":   ["-k31", "-S32", "--prob", "--countsketch-size", "5000"],
    "cs300_pmh_k21_w30_S64": ["-k21", "-w30", "-S64", "--prob", "--countmin-size", "300"],
    "cs100000_bmh_k31_S16": ["-k31", "-S16", "--multiset", "-c", "100000"],          # far more buckets than distinct k-mers: mostly empty
    "cs700_pmh_k31_S32_m3": ["-k31", "-S32", "--prob", "--countsketch-size", "700", "-m", "3"],
}
FILES = ["dup.fa", "g0.fa", "rep.fa", "adv.fa"]


def main():
    work = tempfile.mkdtemp(prefix="d2goldcs")
    paths = []
    for n in FILES:
        dst = os.path.join(work, n); open(dst, "wb").write(gzip.open(os.path.join(INP, n + ".gz"), "rb").read()); paths.append(dst)
    for name, argv in CASES.items():
        out = os.path.join(work, name + ".stk"); cdir = os.path.join(work, "c_" + name); os.makedirs(cdir)
        refbin.run_ref(["sketch", "-p1", "-o", out, "--cache", "--outprefix", cdir] + argv + paths, threads=1)
        cards, sigs = read_stacked(out)
        np.savez_compressed(os.path.join(EXP, name + ".npz"), cards=cards, sigs=sigs, cache_names=np.array(sorted(os.listdir(cdir))))
        print(name, cards, sorted(os.listdir(cdir))[0])
    shutil.rmtree(work)


if __name__ == "__main__":
    main()
