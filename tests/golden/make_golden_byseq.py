#!/usr/bin/env python
"""Golden vectors for `--parse-by-seq` (one sketch per record, src/fastxsketchbyseq.cpp); UNMODIFIED reference binary.
Dev container only (needs oracle/_ref).  Writes tests/golden/inputs/byseq.fa.gz and tests/golden/expected/byseq_*.npz
(cards, sigs, names and -- for the cases listed in CMP -- the all-pairs matrix of the same run)."""
import gzip, os, shutil, sys, tempfile
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__)); ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
from dashing2_b200 import synth  # noqa: E402
import refbin  # noqa: E402
from make_golden import read_stacked  # noqa: E402
INP = os.path.join(HERE, "inputs"); EXP = os.path.join(HERE, "expected")

CASES = {
    "byseq_opmh_k31_S64":      ["-k31", "-S64"],
    "byseq_opmh_k21_w30_S64":  ["-k21", "-w30", "-S64"],
    "byseq_opmh_k15_S16_nocanon": ["-k15", "-S16", "-C"],
    "byseq_fss_k31_S64":       ["-k31", "-S64", "--full-setsketch"],
    "byseq_fss_k21_w30_S32":   ["-k21", "-w30", "-S32", "--full-setsketch"],
    "byseq_bmh_k31_S32":       ["-k31", "-S32", "--multiset"],
    "byseq_pmh_k31_S32":       ["-k31", "-S32", "--prob"],
}
CMP = ("byseq_opmh_k31_S64", "byseq_fss_k31_S64")


def make_input():
    rng = np.random.default_rng(31)
    base = synth._ACGT[rng.integers(0, 4, size=40000)].tobytes()
    def mut(s, rate):
        a = np.frombuffer(s, dtype=np.uint8).copy()
        hit = rng.random(a.size) < rate
        a[hit] = synth._ACGT[rng.integers(0, 4, size=int(hit.sum()))]
        return a.tobytes()
    recs = [("r0", base[:150]), ("r1 long", base[150:3150]), ("r2", base[3150:3190]), ("r3", base[3200:3210] + b"N" + base[3211:3231]),
            ("r4", base[3300:3330]), ("r5", b""), ("r6", base[4000:16000]), ("r7", base[16000:16500]),
            ("r8 lower", base[150:3150].lower()), ("r9 mutated", mut(base[150:3150], 0.02)), ("r10", mut(base[4000:16000], 0.01)),
            ("r11 dup", base[17000:17400] * 3), ("r12 Ns", base[18000:18200] + b"NNN" + base[18203:18500] + b"N" + base[18501:18700]),
            ("r13", b"A" * 120), ("r14", mut(base[16000:16500], 0.05)), ("r15", base[20000:28000])]
    out = b""
    for name, s in recs:
        out += b">" + name.encode() + b"\n"
        for i in range(0, len(s), 70):
            out += s[i:i + 70] + b"\n"
    with gzip.GzipFile(os.path.join(INP, "byseq.fa.gz"), "wb", mtime=0) as f:
        f.write(out)
    return out


def main():
    if refbin.ref_binary() is None:
        sys.exit("reference binary missing: run `make -f oracle/Makefile.ref -j8` first")
    data = make_input()
    work = tempfile.mkdtemp(prefix="d2goldq")
    fa = os.path.join(work, "byseq.fa")
    open(fa, "wb").write(data)
    for name, argv in CASES.items():
        out = os.path.join(work, name + ".stk"); mat = os.path.join(work, name + ".f32")
        extra = ["--binary-output", "--cmpout", mat] if name in CMP else []
        refbin.run_ref(["sketch", "--parse-by-seq", "-p1", "-o", out] + argv + extra + [fa], threads=1)
        cards, sigs = read_stacked(out)
        names = [l.split("\t")[0] for l in open(out + ".names.txt").read().splitlines()[1:]]
        kw = {"mat": np.fromfile(mat, dtype=np.float32)} if extra else {}
        np.savez_compressed(os.path.join(EXP, name + ".npz"), cards=cards, sigs=sigs, names=np.array(names), **kw)
        print(name, cards)
    shutil.rmtree(work)


if __name__ == "__main__":
    main()
