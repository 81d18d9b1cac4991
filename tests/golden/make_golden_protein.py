#!/usr/bin/env python
"""Golden vectors for the protein alphabets (--protein, --protein14, --protein6, --protein8), per file and --parse-by-seq; UNMODIFIED
reference binary.  Dev container only (needs oracle/_ref).  Pins the ORACLE's protein k-mer stream; the GPU encode for these alphabets
is not built yet.  Writes tests/golden/inputs/prot.fa.gz."""
import gzip, os, shutil, sys, tempfile
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__)); ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import refbin  # noqa: E402
from make_golden import read_stacked  # noqa: E402
INP = os.path.join(HERE, "inputs"); EXP = os.path.join(HERE, "expected")
CASES = {
    "prot20_opmh_k7_S64": ["--protein", "-k7", "-S64"],
    "prot20_opmh_k14_S64": ["--protein", "-k14", "-S64"],
    "prot14_opmh_k10_S64": ["--protein14", "-k10", "-S64"],
    "prot6_opmh_k20_S64": ["--protein6", "-k20", "-S64"],
    "prot8_opmh_k12_S64": ["--protein8", "-k12", "-S64"],
    "prot20_opmh_k5_w12_S32": ["--protein", "-k5", "-w12", "-S32"],
    "prot20_fss_k7_S64": ["--protein", "-k7", "-S64", "--full-setsketch"],
}


def make_input():
    rng = np.random.default_rng(61)
    aa = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWY", dtype=np.uint8)
    base = aa[rng.integers(0, 20, size=6000)].tobytes()
    def mut(s, rate):
        a = np.frombuffer(s, dtype=np.uint8).copy(); hit = rng.random(a.size) < rate
        a[hit] = aa[rng.integers(0, 20, size=int(hit.sum()))]; return a.tobytes()
    recs = [("p0", base[:900]), ("p1 mutated", mut(base[:900], 0.05)), ("p2 lower", base[900:1500].lower()),
            ("p3 invalid", base[1500:1700] + b"X" + base[1701:1900] + b"*BZ" + base[1903:2100] + b"OU" + base[2102:2300]),
            ("p4 short", base[2300:2306]), ("p5", base[2400:5000]), ("p6", mut(base[2400:5000], 0.1)), ("p7 empty", b"")]
    out = b"".join(b">" + n.encode() + b"\n" + b"\n".join(s[i:i + 60] for i in range(0, len(s), 60)) + b"\n" for n, s in recs)
    with gzip.GzipFile(os.path.join(INP, "prot.fa.gz"), "wb", mtime=0) as f:
        f.write(out)
    return out


def main():
    data = make_input()
    work = tempfile.mkdtemp(prefix="d2goldpr")
    fa = os.path.join(work, "prot.fa"); open(fa, "wb").write(data)
    for name, argv in CASES.items():
        out = os.path.join(work, name + ".stk")
        refbin.run_ref(["sketch", "-p1", "-o", out] + argv + [fa], threads=1)
        cards, sigs = read_stacked(out)
        out2 = os.path.join(work, name + ".byseq.stk")
        refbin.run_ref(["sketch", "--parse-by-seq", "-p1", "-o", out2] + argv + [fa], threads=1)
        bcards, bsigs = read_stacked(out2)
        np.savez_compressed(os.path.join(EXP, name + ".npz"), cards=cards, sigs=sigs, byseq_cards=bcards, byseq_sigs=bsigs)
        print(name, cards, bcards)
    shutil.rmtree(work)


if __name__ == "__main__":
    main()
