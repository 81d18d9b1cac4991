#!/usr/bin/env python
"""Generate the committed golden vectors by running the UNMODIFIED reference binary.

Run in the dev container only (needs oracle/_ref/dashing2-v4, built from /root/reference by
`make -f oracle/Makefile.ref`):

    python tests/golden/make_golden.py

Inputs (tests/golden/inputs/*.fa.gz, *.npz) are produced from fixed seeds by
dashing2_b200/synth.py and committed; outputs land in tests/golden/expected/.  The parity tests
(tests/test_oracle_vs_golden.py on CPU, tests/test_gpu_*.py on the GPU) read only these files --
never /root/reference and never the reference binary.
"""
import gzip
import json
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from dashing2_b200 import synth  # noqa: E402
import refbin  # noqa: E402

INP = os.path.join(HERE, "inputs")
EXP = os.path.join(HERE, "expected")

# name -> (reference argv after `sketch`, description)
SKETCH_CASES = {
    "opmh_k31_S1024":        ["-k31", "-S1024"],
    "opmh_k31_w51_S512":     ["-k31", "-w51", "-S512"],
    "opmh_k21_S256_nocanon": ["-k21", "-S256", "-C"],
    "opmh_k21_w30_S256_nocanon": ["-k21", "-w30", "-S256", "-C"],
    "opmh_k31_S256_seed17":  ["-k31", "-S256", "--seed", "17"],
    "opmh_k15_S64":          ["-k15", "-S64"],           # empty buckets -> densify matters
    "opmh_k31_S1000":        ["-k31", "-S1000"],         # non power-of-two S
    "fss_k31_S256":          ["-k31", "-S256", "--full-setsketch"],
    "fss_k31_w51_S1024":     ["-k31", "-w51", "-S1024", "--full-setsketch"],
    "bmh_k31_S128":          ["-k31", "-S128", "--multiset", "--cache"],
    "pmh_k31_S128":          ["-k31", "-S128", "--prob", "--cache"],
}
# measures emitted for the opmh_k31_S1024 sketches: name -> extra argv
CMP_CASES = {
    "sim_sym": [],
    "sim_asym": ["--asymmetric-all-pairs"],
    "containment_sym": ["--containment"],
    "symcontainment_sym": ["--symmetric-containment"],
    "mash_sym": ["--mash-distance"],
    "isz_sym": ["--intersection"],
    "usz_sym": ["--union-size"],
}


def make_inputs():
    os.makedirs(INP, exist_ok=True)
    names = []
    for g, seq in synth.family_genomes(6, 30000, seed=11):
        p = os.path.join(INP, f"g{g}.fa.gz")
        with gzip.GzipFile(p, "wb", mtime=0) as f:
            f.write(synth.fasta_bytes(f"g{g}", seq))
        names.append(f"g{g}.fa")
    # duplicated-content genome (k-mer multiplicities > 1 for the multiset sketches)
    for g, seq in synth.family_genomes(1, 20000, seed=12, dup_frac=0.5):
        with gzip.GzipFile(os.path.join(INP, "dup.fa.gz"), "wb", mtime=0) as f:
            f.write(synth.fasta_bytes("dup", seq))
        names.append("dup.fa")
    rng = np.random.default_rng(5)
    s = synth._ACGT[rng.integers(0, 4, size=5000)].tobytes()
    adv = (b">r1 with Ns\n" + s[:1000] + b"NNNN" + s[1000:1500].lower() + b"N" + s[1500:2000] +
           b"\n>r2 short\nACGTACGT\n>r3 shorter than w\n" + s[2000:2040] + b"\n>r4\n" +
           b"\n".join(s[2040 + i:2040 + i + 70] for i in range(0, 2960, 70)) + b"\n>r5 runs\n" +
           b"T" * 40 + b"A" * 40 + s[100:200] + b"\n")
    with gzip.GzipFile(os.path.join(INP, "adv.fa.gz"), "wb", mtime=0) as f:
        f.write(adv)
    names.append("adv.fa")
    fq = b"@q1\n" + s[:150] + b"\n+\n" + b"I" * 150 + b"\n@q2\n" + s[150:260] + b"N" + s[260:300] + b"\n+\n" + b"#" * 151 + b"\n"
    with gzip.GzipFile(os.path.join(INP, "reads.fq.gz"), "wb", mtime=0) as f:
        f.write(fq)
    names.append("reads.fq")
    with open(os.path.join(INP, "order.txt"), "w") as f:
        f.write("\n".join(names) + "\n")
    # synthetic register matrices for the compare / top-k paths
    regs, cards = synth.synthetic_sketches(48, 256, seed=21, n_families=6)
    cards = cards * (1 + np.arange(48) % 5)
    np.savez_compressed(os.path.join(INP, "sk48x256.npz"), regs=regs, cards=cards)
    regs, cards = synth.synthetic_sketches(600, 64, seed=22, n_families=20, p_lo=0.02, p_hi=0.6)
    np.savez_compressed(os.path.join(INP, "sk600x64.npz"), regs=regs, cards=cards)
    return names


def materialise(work):
    names = open(os.path.join(INP, "order.txt")).read().split()
    paths = []
    for n in names:
        dst = os.path.join(work, n)
        with gzip.open(os.path.join(INP, n + ".gz"), "rb") as f, open(dst, "wb") as o:
            shutil.copyfileobj(f, o)
        paths.append(dst)
    return names, paths


def read_stacked(path):
    n, s = (int(x) for x in np.fromfile(path, dtype=np.uint64, count=2))
    d = np.fromfile(path, dtype=np.float64, offset=16)
    return d[:n].copy(), d[n:n + n * s].reshape(n, s).copy()


def main():
    if refbin.ref_binary() is None:
        sys.exit("reference binary missing: run `make -f oracle/Makefile.ref -j8` first")
    os.makedirs(EXP, exist_ok=True)
    make_inputs()
    work = tempfile.mkdtemp(prefix="d2gold")
    names, paths = materialise(work)
    flist = os.path.join(work, "files.txt")
    open(flist, "w").write("\n".join(paths) + "\n")
    manifest = {"reference": "dnbaker/dashing2 v2.1.20 (3906ebde)", "binary": os.path.basename(refbin.ref_binary()),
                "sketch": {}, "cmp": {}}
    for name, argv in SKETCH_CASES.items():
        out = os.path.join(work, name + ".stk")
        extra = ["--outprefix", os.path.join(work, "cache_" + name)] if "--cache" in argv else []
        if extra:
            os.makedirs(extra[1], exist_ok=True)
        refbin.run_ref(["sketch", "-p1", "-F", flist, "-o", out] + argv + extra, threads=1)
        cards, sigs = read_stacked(out)
        np.savez_compressed(os.path.join(EXP, name + ".npz"), cards=cards, sigs=sigs)
        manifest["sketch"][name] = argv
    # --save-kmers (ids) for OPMH
    out = os.path.join(work, "savek.stk")
    refbin.run_ref(["sketch", "-p1", "-F", flist, "-o", out, "-k31", "-S256", "--save-kmers"], threads=1)
    hdr = np.fromfile(out + ".kmer64", dtype=np.uint32, count=4)
    ids = np.fromfile(out + ".kmer64", dtype=np.uint64, offset=24).reshape(len(paths), 256)
    cards, sigs = read_stacked(out)
    np.savez_compressed(os.path.join(EXP, "opmh_k31_S256_savekmers.npz"), cards=cards, sigs=sigs, ids=ids, hdr=hdr)
    # sketch + cmp in one go (config-1 shape): binary matrices for each measure (sigs get densified in place)
    for name, argv in CMP_CASES.items():
        out = os.path.join(work, name + ".stk")
        mat = os.path.join(work, name + ".f32")
        refbin.run_ref(["sketch", "-p1", "-F", flist, "-o", out, "-k31", "-S1024", "--binary-output", "--cmpout", mat] + argv, threads=1)
        np.save(os.path.join(EXP, "cmp_opmh_k31_S1024_" + name + ".npy"), np.fromfile(mat, dtype=np.float32))
        manifest["cmp"][name] = argv
    # low-S sketch whose empty buckets force densification before compare
    out = os.path.join(work, "dens.stk"); mat = os.path.join(work, "dens.f32")
    refbin.run_ref(["sketch", "-p1", "-F", flist, "-o", out, "-k15", "-S64", "--binary-output", "--cmpout", mat], threads=1)
    cards, sigs = read_stacked(out)
    np.savez_compressed(os.path.join(EXP, "opmh_k15_S64_densified.npz"), cards=cards, sigs=sigs,
                        mat=np.fromfile(mat, dtype=np.float32))
    # text outputs (phylip + default table) for the format writer
    for tag, argv in (("phylip", ["--phylip"]), ("table", [])):
        mat = os.path.join(work, tag + ".txt")
        refbin.run_ref(["sketch", "-p1", "-F", flist, "-k31", "-S1024", "--cmpout", mat] + argv, threads=1, cwd=work)
        txt = open(mat).read().replace(work + "/", "")
        open(os.path.join(EXP, "cmp_opmh_k31_S1024_" + tag + ".txt"), "w").write(txt)
    # presketched compare on synthetic registers: symmetric / panel / measures / equality kinds
    z = np.load(os.path.join(INP, "sk48x256.npz"))
    for suffix, kinds in ((".ss", ("sim_sym", "sim_asym", "containment_sym", "symcontainment_sym", "mash_sym", "isz_sym", "usz_sym")),
                          (".bmh", ("sim_sym", "containment_sym", "isz_sym", "mash_sym", "usz_sym", "symcontainment_sym"))):
        stk = os.path.join(work, "sk48" + suffix)
        synth.write_stacked(stk, z["regs"], z["cards"], names=[f"s{i}" for i in range(48)])
        for kind in kinds:
            mat = os.path.join(work, "sk48_" + kind + suffix + ".f32")
            refbin.run_ref(["cmp", "--presketched", "-p1", "--binary-output", "--cmpout", mat, stk] + CMP_CASES[kind], threads=1)
            np.save(os.path.join(EXP, f"cmp_sk48{suffix}_{kind}.npy"), np.fromfile(mat, dtype=np.float32))
    # top-k CSR (LSH path), -p1 (deterministic)
    z = np.load(os.path.join(INP, "sk600x64.npz"))
    stk = os.path.join(work, "sk600.ss")
    synth.write_stacked(stk, z["regs"], z["cards"], names=[f"s{i}" for i in range(600)])
    for K in (5, 32):
        mat = os.path.join(work, f"sk600_top{K}.csr")
        refbin.run_ref(["cmp", "--presketched", "-p1", "--binary-output", "--topk", str(K), "--cmpout", mat, stk], threads=1)
        shutil.copy(mat, os.path.join(EXP, f"topk{K}_sk600.csr"))
    json.dump(manifest, open(os.path.join(EXP, "manifest.json"), "w"), indent=1)
    shutil.rmtree(work)
    print("golden vectors written to", EXP)


if __name__ == "__main__":
    main()
