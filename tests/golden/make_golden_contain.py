#!/usr/bin/env python
"""Golden vectors for `dashing2 contain` (src/contain_main.cpp:133-301): coverage of every reference's sampled k-mers (a FILE.kmer64 written by
`sketch --save-kmers`) by the k-mer stream of each query file, and the mean multiplicity of the covered k-mers.  UNMODIFIED reference binary.
Dev container only (needs oracle/_ref).  Binary output (-b): u64 n_refs, u64 n_queries, f32 coverage[nq][n_refs], f32 mean depth[nq][n_refs]."""
import gzip, os, shutil, sys, tempfile
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__)); ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import refbin  # noqa: E402
INP = os.path.join(HERE, "inputs"); EXP = os.path.join(HERE, "expected")

CASES = {   # database sketch argv
    "contain_opmh_k31_S64": ["-k31", "-S64"],
    "contain_opmh_k21_w30_S32_seed5": ["-k21", "-w30", "-S32", "--seed", "5"],
    "contain_fss_k31_S64": ["-k31", "-S64", "--full-setsketch"],
}
REFS = ["g0.fa", "g1.fa", "dup.fa", "adv.fa"]
QUERIES = ["g0.fa", "rep.fa", "reads.fq", "adv.fa", "dup.fa"]


def main():
    work = tempfile.mkdtemp(prefix="d2goldc")
    paths = {}
    for n in sorted(set(REFS + QUERIES)):
        dst = os.path.join(work, n)
        open(dst, "wb").write(gzip.open(os.path.join(INP, n + ".gz"), "rb").read())
        paths[n] = dst
    for name, argv in CASES.items():
        db = os.path.join(work, name + ".stk")
        refbin.run_ref(["sketch", "-p1", "--save-kmers", "-o", db] + argv + [paths[n] for n in REFS], threads=1)
        outb = os.path.join(work, name + ".bin")
        r = refbin.run_ref(["contain", "-b", "-p1", "-o", outb, db + ".kmer64"] + [paths[n] for n in QUERIES], threads=1, check=False)
        print(name, "rc", r.returncode, r.stderr.decode()[-200:])
        raw = open(outb, "rb").read()
        nref, nq = (int(x) for x in np.frombuffer(raw, np.uint64, 2))
        mat = np.frombuffer(raw, np.float32, 2 * nref * nq, 16).reshape(2, nq, nref)
        hdr = np.fromfile(db + ".kmer64", dtype=np.uint32, count=4)
        ids = np.fromfile(db + ".kmer64", dtype=np.uint64, offset=24).reshape(len(REFS), -1)
        np.savez_compressed(os.path.join(EXP, name + ".npz"), coverage=mat[0], depth=mat[1], hdr=hdr, ids=ids, seed=np.fromfile(db + ".kmer64", dtype=np.uint64, count=3)[2])
        print(mat[0]); print(mat[1])
        if name == "contain_opmh_k31_S64":   # the text form (4 references: below the width of the reference's SIMD formatting blocks)
            outt = os.path.join(work, name + ".txt")
            refbin.run_ref(["contain", "-p1", "-o", outt, db + ".kmer64"] + [paths[n] for n in QUERIES], threads=1)
            txt = open(outt).read().replace(work + "/", "")
            open(os.path.join(EXP, name + ".txt"), "w").write(txt)
            print(txt)
    shutil.rmtree(work)


if __name__ == "__main__":
    main()
