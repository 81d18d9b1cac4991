#!/usr/bin/env python
"""Golden vectors for `--filterset PATH[:x]` (src/d2.cpp:45-98, src/filterset.h, lfunc in src/fastxsketch.cpp:385-398): hashed k-mers found in
the filter set never reach the sketch.  PATH alone = FASTX whose (maskfn'd) k-mers form the set; PATH:B (a colon followed by anything but
K/k) = a raw file of 64-bit hashed values.  UNMODIFIED reference binary, -p1.  Dev container only (needs oracle/_ref)."""
import gzip, os, shutil, sys, tempfile
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__)); ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import refbin  # noqa: E402
import oracle_lib as O  # noqa: E402
from make_golden import read_stacked  # noqa: E402
INP = os.path.join(HERE, "inputs"); EXP = os.path.join(HERE, "expected")

# (case, sketch argv, filter source): the filter is dup.fa (shares most of its k-mers with g0 / g1) or a raw k-mer file
CASES = {
    "fs_opmh_k31_S128": (["-k31", "-S128"], "fasta"),
    "fs_opmh_k21_w30_S64": (["-k21", "-w30", "-S64"], "fasta"),
    "fs_fss_k31_S64": (["-k31", "-S64", "--full-setsketch"], "fasta"),
    "fs_bmh_k31_S32": (["-k31", "-S32", "--multiset"], "fasta"),
    "fs_opmh_k31_S128_raw": (["-k31", "-S128"], "raw"),
    "fs_opmh_k40_S64": (["-k40", "-S64"], "fasta"),
}
FILES = ["g0.fa", "g1.fa", "dup.fa", "adv.fa"]


def main():
    work = tempfile.mkdtemp(prefix="d2goldfs")
    paths = []
    for n in FILES:
        dst = os.path.join(work, n)
        open(dst, "wb").write(gzip.open(os.path.join(INP, n + ".gz"), "rb").read())
        paths.append(dst)
    flist = os.path.join(work, "files.txt"); open(flist, "w").write("\n".join(paths) + "\n")
    fsfa = os.path.join(work, "dup.fa")
    # raw k-mer file: the hashed 31-mers of g1.fa, every third one
    hv = np.concatenate([O.hash_stream(r, 31) for r in O.read_fastx(os.path.join(INP, "g1.fa.gz"))])[::3].copy()
    raw = os.path.join(work, "kmers.u64"); hv.tofile(raw)
    np.save(os.path.join(INP, "filterset_raw_kmers.npy"), hv)
    for name, (argv, src) in CASES.items():
        out = os.path.join(work, name + ".stk")
        fs = fsfa if src == "fasta" else raw + ":B"
        r = refbin.run_ref(["sketch", "-p1", "-F", flist, "-o", out, "--filterset", fs] + argv, threads=1, check=False)
        if r.returncode != 0:
            print(name, "reference binary failed with return code", r.returncode); continue
        cards, sigs = read_stacked(out)
        np.savez_compressed(os.path.join(EXP, name + ".npz"), cards=cards, sigs=sigs)
        print(name, cards, r.stderr.decode()[-200:].replace("\n", " | "))
    shutil.rmtree(work)


if __name__ == "__main__":
    main()
