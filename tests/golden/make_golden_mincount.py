#!/usr/bin/env python
"""Golden vectors for One-Permutation MinHash with `--count-threshold c` (LazyOnePermSetSketch::update with mincount,
src/oph.h:188-205); UNMODIFIED reference binary, per-file and --parse-by-seq.  Dev container only (needs oracle/_ref).
Inputs: the committed fixtures plus tests/golden/inputs/rep.fa.gz (a genome whose thirds overlap, so many k-mers occur 2-3 times)."""
import gzip, os, shutil, sys, tempfile
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__)); ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
from dashing2_b200 import synth  # noqa: E402
import refbin  # noqa: E402
from make_golden import read_stacked  # noqa: E402
INP = os.path.join(HERE, "inputs"); EXP = os.path.join(HERE, "expected")

CASES = {
    "mincount2_opmh_k31_S64":     ["-k31", "-S64", "-m", "2"],
    "mincount3_opmh_k21_S128":    ["-k21", "-S128", "--count-threshold", "3"],
    "mincount2_opmh_k21_w30_S64": ["-k21", "-w30", "-S64", "-m", "2"],
}
FILES = ["rep.fa", "dup.fa", "g0.fa", "adv.fa"]


def make_input():
    rng = np.random.default_rng(41)
    a = synth._ACGT[rng.integers(0, 4, size=6000)].tobytes()
    recs = [("x1", a[:4000]), ("x2", a[2000:6000]), ("x3 third copy", a[3000:5000]), ("x4", a[:300] + b"N" + a[300:600])]
    out = b"".join(b">" + n.encode() + b"\n" + s + b"\n" for n, s in recs)
    with gzip.GzipFile(os.path.join(INP, "rep.fa.gz"), "wb", mtime=0) as f:
        f.write(out)


def main():
    if refbin.ref_binary() is None:
        sys.exit("reference binary missing: run `make -f oracle/Makefile.ref -j8` first")
    make_input()
    work = tempfile.mkdtemp(prefix="d2goldm")
    paths = []
    for n in FILES:
        dst = os.path.join(work, n)
        open(dst, "wb").write(gzip.open(os.path.join(INP, n + ".gz"), "rb").read())
        paths.append(dst)
    flist = os.path.join(work, "files.txt"); open(flist, "w").write("\n".join(paths) + "\n")
    for name, argv in CASES.items():
        out = os.path.join(work, name + ".stk")
        refbin.run_ref(["sketch", "-p1", "-F", flist, "-o", out] + argv, threads=1)
        cards, sigs = read_stacked(out)
        out2 = os.path.join(work, name + ".byseq.stk")
        refbin.run_ref(["sketch", "--parse-by-seq", "-p1", "-o", out2] + argv + [paths[0]], threads=1)
        bcards, bsigs = read_stacked(out2)
        np.savez_compressed(os.path.join(EXP, name + ".npz"), cards=cards, sigs=sigs, byseq_cards=bcards, byseq_sigs=bsigs)
        print(name, cards, (sigs != 0).sum(1), bcards)
    shutil.rmtree(work)


if __name__ == "__main__":
    main()
