#!/usr/bin/env python
"""Golden vectors for `--save-kmercounts` with the one-permutation sketch (multiplicity of each register's minimum, src/oph.h:206-209;
FILE.kmercounts.f64 holds float32[n][S] despite its name, src/fastxsketch.h kmercounts_); UNMODIFIED reference binary.  Dev container only.
Pins the ORACLE's count vector; the GPU library does not produce counts yet."""
import gzip, os, shutil, sys, tempfile
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__)); ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import refbin  # noqa: E402
INP = os.path.join(HERE, "inputs"); EXP = os.path.join(HERE, "expected")
FILES = ["dup.fa", "g0.fa", "rep.fa", "adv.fa"]


def main():
    work = tempfile.mkdtemp(prefix="d2goldk")
    paths = []
    for n in FILES:
        dst = os.path.join(work, n); open(dst, "wb").write(gzip.open(os.path.join(INP, n + ".gz"), "rb").read()); paths.append(dst)
    for name, argv, S in (("kmercounts_opmh_k31_S64", ["-k31", "-S64"], 64), ("kmercounts_opmh_k21_w30_S128", ["-k21", "-w30", "-S128"], 128)):
        out = os.path.join(work, name + ".stk")
        refbin.run_ref(["sketch", "-p1", "-o", out, "--save-kmers", "--save-kmercounts"] + argv + paths, threads=1)
        counts = np.fromfile(out + ".kmercounts.f64", dtype=np.float32).reshape(len(paths), S)
        np.savez_compressed(os.path.join(EXP, name + ".npz"), counts=counts)
        print(name, counts.sum(1), counts.max())
    # the other sketch types: the count kept with a register is the multiplicity (Full SetSketch, setsketch.h:405-406) or the weight
    # (BagMinHash / ProbMinHash) of the k-mer that owns it
    for name, argv, S in (("kmercounts_fss_k31_S64", ["-k31", "-S64", "--full-setsketch"], 64), ("kmercounts_bmh_k31_S32", ["-k31", "-S32", "--multiset", "--cache"], 32),
                          ("kmercounts_pmh_k31_S32", ["-k31", "-S32", "--prob", "--cache"], 32)):
        out = os.path.join(work, name + ".stk"); cdir = os.path.join(work, "c_" + name); os.makedirs(cdir)
        refbin.run_ref(["sketch", "-p1", "-o", out, "--save-kmers", "--save-kmercounts", "--outprefix", cdir] + argv + paths, threads=1)
        counts = np.fromfile(out + ".kmercounts.f64", dtype=np.float32).reshape(len(paths), S)
        np.savez_compressed(os.path.join(EXP, name + ".npz"), counts=counts)
        print(name, counts.sum(1), counts.max())
    shutil.rmtree(work)


if __name__ == "__main__":
    main()
