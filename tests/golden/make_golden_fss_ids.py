#!/usr/bin/env python
"""Golden vectors for `--save-kmers` with the Full SetSketch (ids_[idx] = id where CSetSketch::update lowers a register,
src/setsketch.h:400-404); UNMODIFIED reference binary.  Dev container only (needs oracle/_ref)."""
import gzip, os, shutil, sys, tempfile
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__)); ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import refbin  # noqa: E402
from make_golden import read_stacked  # noqa: E402
INP = os.path.join(HERE, "inputs"); EXP = os.path.join(HERE, "expected")
CASES = {
    "ids_fss_k31_S256": ["-k31", "-S256"],
    "ids_fss_k21_w30_S64_seed5": ["-k21", "-w30", "-S64", "--seed", "5"],
    "ids_fss_k15_S1024": ["-k15", "-S1024"],
}
FILES = ["dup.fa", "g0.fa", "g1.fa", "adv.fa", "reads.fq"]


def main():
    work = tempfile.mkdtemp(prefix="d2goldf")
    paths = []
    for n in FILES:
        dst = os.path.join(work, n); open(dst, "wb").write(gzip.open(os.path.join(INP, n + ".gz"), "rb").read()); paths.append(dst)
    for name, argv in CASES.items():
        out = os.path.join(work, name + ".stk")
        S = int([a for a in argv if a.startswith("-S")][0][2:])
        refbin.run_ref(["sketch", "-p1", "-o", out, "--save-kmers", "--full-setsketch"] + argv + paths, threads=1)
        cards, sigs = read_stacked(out)
        hdr = np.fromfile(out + ".kmer64", dtype=np.uint32, count=4)
        ids = np.fromfile(out + ".kmer64", dtype=np.uint64, offset=24).reshape(len(paths), S)
        np.savez_compressed(os.path.join(EXP, name + ".npz"), cards=cards, sigs=sigs, ids=ids, hdr=hdr)
        print(name, hdr, (ids == 0).sum(1))
    shutil.rmtree(work)


if __name__ == "__main__":
    main()
