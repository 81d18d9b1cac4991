"""The drop-in front-end (dashing2_b200/bin/dashing2-gpu) against reference-binary goldens: same argv as
`dashing2 sketch|cmp`, byte-identical stacked files / binary matrices / text tables."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, expected, GOLD

pytestmark = pytest.mark.gpu
EXE = os.path.join(ROOT, "dashing2_b200", "bin", "dashing2-gpu")


def run(args, cwd=None, env=None):
    r = subprocess.run([EXE] + args, cwd=cwd, capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr
    return r


def read_stacked(path):
    n, s = (int(x) for x in np.fromfile(path, dtype=np.uint64, count=2))
    d = np.fromfile(path, dtype=np.float64, offset=16)
    return d[:n], d[n:].reshape(n, s)


@pytest.mark.parametrize("case,argv", [("opmh_k31_S1024", ["-k31", "-S1024"]), ("opmh_k31_w51_S512", ["-k31", "-w51", "-S512"]),
                                       ("opmh_k31_S256_seed17", ["-k31", "-S256", "--seed", "17"]),
                                       ("fss_k31_w51_S1024", ["-k31", "-w51", "-S1024", "--full-setsketch"])])
def test_sketch_stacked_file(case, argv, golden_inputs, tmp_path):
    names, paths = golden_inputs
    flist = tmp_path / "files.txt"; flist.write_text("\n".join(paths) + "\n")
    out = str(tmp_path / "out.stk")
    run(["sketch", "-p4", "-F", str(flist), "-o", out] + argv)
    cards, sigs = read_stacked(out)
    z = np.load(expected(case + ".npz"))
    assert np.array_equal(sigs.view(np.uint64), z["sigs"].view(np.uint64))
    if case.startswith("opmh"):
        assert np.array_equal(cards, z["cards"])
    else:
        np.testing.assert_allclose(cards, z["cards"], rtol=1e-12)
    lines = open(out + ".names.txt").read().splitlines()
    assert lines[0] == "#Name\tCardinality" and [l.split("\t")[0] for l in lines[1:]] == paths
    if case.startswith("opmh"):
        assert [l.split("\t")[1] for l in lines[1:]] == ["%0.24g" % c for c in z["cards"]]


def test_sketch_cmpout_binary_text_and_cache(golden_inputs, tmp_path):
    names, paths = golden_inputs
    flist = tmp_path / "files.txt"; flist.write_text("\n".join(paths) + "\n")
    work = os.path.dirname(paths[0])
    for kind, argv in (("sim_sym", []), ("sim_asym", ["--asymmetric-all-pairs"]), ("containment_sym", ["--containment"]),
                       ("mash_sym", ["--mash-distance"]), ("usz_sym", ["--union-size"])):
        mat = str(tmp_path / (kind + ".f32"))
        run(["sketch", "-F", str(flist), "-k31", "-S1024", "--binary-output", "--cmpout", mat] + argv)
        exp = np.load(expected(f"cmp_opmh_k31_S1024_{kind}.npy"))
        assert np.array_equal(np.fromfile(mat, dtype=np.float32).view(np.uint32), exp.view(np.uint32)), kind
    for tag, argv in (("phylip", ["--phylip"]), ("table", [])):
        mat = str(tmp_path / (tag + ".txt"))
        run(["sketch", "-F", str(flist), "-k31", "-S1024", "--cmpout", mat] + argv)
        got = open(mat).read().replace(work + "/", "")
        assert got == open(expected(f"cmp_opmh_k31_S1024_{tag}.txt")).read(), tag
    # --cache writes reference-named per-file sketches that a second run (and `cmp --cache`) reloads
    cdir = tmp_path / "cache"; cdir.mkdir()
    run(["sketch", "-F", str(flist), "-k31", "-S1024", "--cache", "--outprefix", str(cdir)])
    f0 = cdir / (os.path.basename(paths[0]) + ".rc_canon.sketchsize1024.k31.SetSpace.DNA.opss")
    assert f0.exists() and f0.stat().st_size == 8 + 1024 * 8
    z = np.load(expected("opmh_k31_S1024.npz"))
    d = np.fromfile(f0, dtype=np.float64)
    assert d[0] == z["cards"][0] and np.array_equal(d[1:], z["sigs"][0])
    mat = str(tmp_path / "fromcache.f32")
    run(["cmp", "--cache", "--outprefix", str(cdir), "-k31", "-S1024", "--binary-output", "--cmpout", mat, "-F", str(flist)])
    exp = np.load(expected("cmp_opmh_k31_S1024_sim_sym.npy"))
    assert np.array_equal(np.fromfile(mat, dtype=np.float32), exp)


def test_cmp_presketched_and_panel(tmp_path):
    from dashing2_b200 import synth
    import oracle_lib as O
    z = np.load(os.path.join(GOLD, "inputs", "sk48x256.npz"))
    for suffix, cmp_kind in ((".ss", 0), (".bmh", 1)):
        stk = str(tmp_path / ("sk48" + suffix))
        synth.write_stacked(stk, z["regs"], z["cards"], names=[f"s{i}" for i in range(48)])
        for kind, argv in (("sim_sym", []), ("symcontainment_sym", ["--symmetric-containment"]), ("isz_sym", ["--intersection"])):
            mat = str(tmp_path / "m.f32")
            run(["cmp", "--presketched", "--binary-output", "--cmpout", mat, stk] + argv)
            exp = np.load(expected(f"cmp_sk48{suffix}_{kind}.npy"))
            assert np.array_equal(np.fromfile(mat, dtype=np.float32).view(np.uint32), exp.view(np.uint32)), (suffix, kind)


def test_cmp_fastcmp_and_bbit_sigs(tmp_path):
    """`cmp --presketched --fastcmp N [--bbit-sigs]`: byte-identical matrices to the reference binary, and the same fitted
    (a, b) on stderr."""
    import json
    from dashing2_b200 import synth
    z = np.load(os.path.join(GOLD, "inputs", "sk48x256.npz"))
    stk = str(tmp_path / "sk48.ss")
    synth.write_stacked(stk, z["regs"], z["cards"], names=[f"s{i}" for i in range(48)])
    ab = json.load(open(expected("cmpc_fitted_ab.json")))
    for fd in ("1", "2", "4"):
        for bbit in (False, True):
            for kind, argv in (("sim_sym", []), ("mash_sym", ["--mash-distance"]), ("containment_sym", ["--containment"]), ("sim_asym", ["--asymmetric-all-pairs"])):
                mat = str(tmp_path / "m.f32")
                r = run(["cmp", "--presketched", "--binary-output", "--cmpout", mat, "--fastcmp", fd] + (["--bbit-sigs"] if bbit else []) + argv + [stk])
                exp = np.load(expected(f"cmpc_sk48_fd{fd}_{'bbit' if bbit else 'ss'}_{kind}.npy"))
                assert np.array_equal(np.fromfile(mat, dtype=np.float32).view(np.uint32), exp.view(np.uint32)), (fd, bbit, kind)
                if not bbit:
                    assert f"a = {ab['fd' + fd][0]} and b = {ab['fd' + fd][1]}" in r.stderr


def test_multiset_and_prob_cache_files(golden_inputs, tmp_path):
    """--multiset / --prob --cache: reference-named .bmh / .pmh cache files with reference bytes."""
    names, paths = golden_inputs
    flist = tmp_path / "files.txt"; flist.write_text("\n".join(paths) + "\n")
    for flag, tag, space, suf in (("--multiset", "bmh", "MultisetSpace", ".bmh"), ("--prob", "pmh", "ProbsetSpace", ".pmh")):
        cdir = tmp_path / tag; cdir.mkdir()
        out = str(tmp_path / (tag + ".stk"))
        run(["sketch", "-F", str(flist), "-k31", "-S128", flag, "--cache", "--outprefix", str(cdir), "-o", out])
        z = np.load(expected(f"{tag}_k31_S128.npz"))
        cards, sigs = read_stacked(out)
        assert np.array_equal(sigs.view(np.uint64), z["sigs"].view(np.uint64)) and np.array_equal(cards, z["cards"])
        f0 = cdir / (os.path.basename(paths[0]) + f".rc_canon.sketchsize128.k31.ExactCounting.{space}.DNA{suf}")
        d = np.fromfile(f0, dtype=np.float64)
        assert d[0] == z["cards"][0] and np.array_equal(d[1:], z["sigs"][0])


@pytest.mark.parametrize("case,argv", [("byseq_opmh_k31_S64", ["-k31", "-S64"]), ("byseq_opmh_k21_w30_S64", ["-k21", "-w30", "-S64"]),
                                       ("byseq_fss_k31_S64", ["-k31", "-S64", "--full-setsketch"]), ("byseq_bmh_k31_S32", ["-k31", "-S32", "--multiset"]),
                                       ("byseq_pmh_k31_S32", ["-k31", "-S32", "--prob"])])
def test_parse_by_seq_stacked_file_names_and_matrix(case, argv, tmp_path):
    """`sketch --parse-by-seq`: one sketch per record, record names, exact cardinalities below 10 * S, and the all-pairs matrix
    of the same run -- the stacked file and the matrix byte for byte as the reference binary wrote them."""
    import gzip
    fa = str(tmp_path / "byseq.fa")
    open(fa, "wb").write(gzip.open(os.path.join(GOLD, "inputs", "byseq.fa.gz"), "rb").read())
    z = np.load(expected(case + ".npz"))
    out = str(tmp_path / "out.stk"); mat = str(tmp_path / "out.f32")
    extra = ["--binary-output", "--cmpout", mat] if "mat" in z.files else []
    run(["sketch", "--parse-by-seq", "-p2", "-o", out] + argv + extra + [fa])
    cards, sigs = read_stacked(out)
    assert np.array_equal(sigs.view(np.uint64), z["sigs"].view(np.uint64))
    if "fss" in case:
        np.testing.assert_allclose(cards, z["cards"], rtol=1e-12)
    else:
        assert np.array_equal(cards, z["cards"])
    lines = open(out + ".names.txt").read().splitlines()
    assert [l.split("\t")[0] for l in lines[1:]] == list(z["names"])
    if extra:   # similarity does not involve the cardinalities
        assert np.array_equal(np.fromfile(mat, dtype=np.float32).view(np.uint32), z["mat"].view(np.uint32))


def test_count_threshold_per_file_and_by_seq(tmp_path):
    """`sketch -m 2`: reference-named cache files (.ct_threshold2) and stacked file; `--parse-by-seq -m 2` reproduces the reference,
    whose per-record one-permutation sketches ignore the threshold."""
    import gzip
    paths = []
    for f in ("rep.fa", "dup.fa", "g0.fa", "adv.fa"):
        p = str(tmp_path / f); open(p, "wb").write(gzip.open(os.path.join(GOLD, "inputs", f + ".gz"), "rb").read()); paths.append(p)
    z = np.load(expected("mincount2_opmh_k31_S64.npz"))
    out = str(tmp_path / "out.stk"); cdir = tmp_path / "cache"; cdir.mkdir()
    run(["sketch", "-k31", "-S64", "-m", "2", "-o", out, "--cache", "--outprefix", str(cdir)] + paths)
    cards, sigs = read_stacked(out)
    assert np.array_equal(sigs.view(np.uint64), z["sigs"].view(np.uint64)) and np.array_equal(cards, z["cards"])
    assert (cdir / "rep.fa.rc_canon.sketchsize64.k31.ct_threshold2.SetSpace.DNA.opss").exists()
    out2 = str(tmp_path / "byseq.stk")
    run(["sketch", "--parse-by-seq", "-k31", "-S64", "--count-threshold", "2", "-o", out2, paths[0]])
    cards, sigs = read_stacked(out2)
    assert np.array_equal(sigs.view(np.uint64), z["byseq_sigs"].view(np.uint64)) and np.array_equal(cards, z["byseq_cards"])


@pytest.mark.parametrize("share", [True, False])
def test_gpus_option_shards_files_and_rows(share, golden_inputs, tmp_path, monkeypatch):
    """--gpus N (front-end only): batches of files go to whichever device is free, every device computes a range of output rows /
    neighbour lists; outputs stay byte-identical to the reference goldens.  share=True runs three contexts on one device
    (D2G_SHARE_DEVICE), share=False two contexts on two devices."""
    from dashing2_b200 import synth
    if share:
        monkeypatch.setenv("D2G_SHARE_DEVICE", "1")
    else:
        import torch
        if torch.cuda.device_count() < 2:
            pytest.skip("needs 2 GPUs")
    ng = "3" if share else "2"
    names, paths = golden_inputs
    flist = tmp_path / "files.txt"; flist.write_text("\n".join(paths) + "\n")
    work = os.path.dirname(paths[0])
    out = str(tmp_path / "out.stk"); mat = str(tmp_path / "m.f32")
    run(["sketch", "--gpus", ng, "-p6", "-F", str(flist), "-k31", "-S1024", "-o", out, "--binary-output", "--cmpout", mat])
    exp = np.load(expected("cmp_opmh_k31_S1024_sim_sym.npy"))
    assert np.array_equal(np.fromfile(mat, dtype=np.float32).view(np.uint32), exp.view(np.uint32))
    cards, _ = read_stacked(out)
    assert np.array_equal(cards, np.load(expected("opmh_k31_S1024.npz"))["cards"])
    for tag, argv in (("phylip", ["--phylip"]), ("table", [])):
        txt = str(tmp_path / (tag + ".txt"))
        run(["sketch", "--gpus", ng, "-F", str(flist), "-k31", "-S1024", "--cmpout", txt] + argv)
        assert open(txt).read().replace(work + "/", "") == open(expected(f"cmp_opmh_k31_S1024_{tag}.txt")).read(), tag
    mat = str(tmp_path / "asym.f32")
    run(["sketch", "--gpus", ng, "-F", str(flist), "-k31", "-S1024", "--binary-output", "--asymmetric-all-pairs", "--cmpout", mat])
    exp = np.load(expected("cmp_opmh_k31_S1024_sim_asym.npy"))
    assert np.array_equal(np.fromfile(mat, dtype=np.float32).view(np.uint32), exp.view(np.uint32))
    z = np.load(os.path.join(GOLD, "inputs", "sk600x64.npz"))
    stk = str(tmp_path / "sk600.ss")
    synth.write_stacked(stk, z["regs"], z["cards"], names=[f"s{i}" for i in range(600)])
    csr = str(tmp_path / "top5.csr")
    run(["cmp", "--gpus", ng, "--presketched", "--binary-output", "--topk", "5", "--cmpout", csr, stk])
    assert open(csr, "rb").read() == open(expected("topk5_sk600.csr"), "rb").read()
    z = np.load(os.path.join(GOLD, "inputs", "sk48x256.npz"))
    stk = str(tmp_path / "sk48.ss")
    synth.write_stacked(stk, z["regs"], z["cards"], names=[f"s{i}" for i in range(48)])
    run(["cmp", "--gpus", ng, "--presketched", "--binary-output", "--cmpout", mat, stk])
    assert np.array_equal(np.fromfile(mat, dtype=np.float32).view(np.uint32), np.load(expected("cmp_sk48.ss_sim_sym.npy")).view(np.uint32))


@pytest.mark.parametrize("case,argv", [("panel_opmh_k31_S1024_sim", ["-k31", "-S1024"]), ("panel_opmh_k31_S1024_containment", ["-k31", "-S1024", "--containment"]),
                                       ("panel_fss_k31_S256_mash", ["-k31", "-S256", "--full-setsketch", "--mash-distance"])])
def test_panel_from_fasta_lists(case, argv, golden_inputs, tmp_path):
    """`sketch -F refs -Q queries --cmpout`: |F| x |Q| matrix, binary and text, byte for byte as the reference binary wrote them."""
    names, paths = golden_inputs
    work = os.path.dirname(paths[0])
    ff = tmp_path / "refs.txt"; ff.write_text("\n".join(paths[:4]) + "\n")
    qf = tmp_path / "queries.txt"; qf.write_text("\n".join(paths[4:]) + "\n")
    mat = str(tmp_path / "p.f32"); txt = str(tmp_path / "p.txt")
    run(["sketch", "-F", str(ff), "-Q", str(qf), "--binary-output", "--cmpout", mat] + argv)
    assert np.array_equal(np.fromfile(mat, dtype=np.float32).view(np.uint32), np.load(expected(case + ".npy")).view(np.uint32))
    run(["sketch", "-F", str(ff), "-Q", str(qf), "--cmpout", txt] + argv)
    assert open(txt).read().replace(work + "/", "") == open(expected(case + ".txt")).read()
    run(["sketch", "--gpus", "2", "-F", str(ff), "-Q", str(qf), "--cmpout", txt] + argv, env=dict(os.environ, D2G_SHARE_DEVICE="1"))
    assert open(txt).read().replace(work + "/", "") == open(expected(case + ".txt")).read()


def test_countsketch_size_cache_names_and_registers(tmp_path):
    """`sketch --prob --countsketch-size 700 -m 3 --cache`: reference file names (.ct_threshold3.CountMinCounting700.) and bytes."""
    import gzip
    paths = []
    for f in ("dup.fa", "g0.fa", "rep.fa", "adv.fa"):
        p = str(tmp_path / f); open(p, "wb").write(gzip.open(os.path.join(GOLD, "inputs", f + ".gz"), "rb").read()); paths.append(p)
    z = np.load(expected("cs700_pmh_k31_S32_m3.npz"))
    out = str(tmp_path / "out.stk"); cdir = tmp_path / "cache"; cdir.mkdir()
    run(["sketch", "-k31", "-S32", "--prob", "--countsketch-size", "700", "-m", "3", "-o", out, "--cache", "--outprefix", str(cdir)] + paths)
    cards, sigs = read_stacked(out)
    assert np.array_equal(sigs.view(np.uint64), z["sigs"].view(np.uint64)) and np.array_equal(cards, z["cards"])
    assert sorted(os.listdir(cdir)) == list(z["cache_names"])
    z = np.load(expected("cs100000_bmh_k31_S16.npz"))
    run(["sketch", "-k31", "-S16", "--multiset", "-c", "100000", "-o", out] + paths)
    cards, sigs = read_stacked(out)
    assert np.array_equal(sigs.view(np.uint64), z["sigs"].view(np.uint64)) and np.array_equal(cards, z["cards"])


def test_save_kmers_with_counting_sketches(tmp_path):
    """`sketch --multiset --save-kmers -o FILE`: FILE.kmer64 (header + ids) as the reference binary wrote it."""
    import gzip
    paths = []
    for f in ("dup.fa", "g0.fa", "rep.fa", "adv.fa", "reads.fq"):
        p = str(tmp_path / f); open(p, "wb").write(gzip.open(os.path.join(GOLD, "inputs", f + ".gz"), "rb").read()); paths.append(p)
    for case, argv in (("ids_bmh_k31_S64", ["-k31", "-S64", "--multiset"]), ("ids_pmh_k21_w30_S32_seed5", ["-k21", "-w30", "-S32", "--prob", "--seed", "5"])):
        assert case  # the Full SetSketch .kmer64 file is covered by tests/test_gpu_parity.py::test_fss_save_kmers_ids_match_reference_golden
        z = np.load(expected(case + ".npz"))
        out = str(tmp_path / (case + ".stk"))
        run(["sketch", "--save-kmers", "-o", out] + argv + paths)
        S = z["sigs"].shape[1]
        assert np.array_equal(np.fromfile(out + ".kmer64", dtype=np.uint32, count=4), z["hdr"])
        assert np.array_equal(np.fromfile(out + ".kmer64", dtype=np.uint64, offset=24).reshape(len(paths), S), z["ids"])
        cards, sigs = read_stacked(out)
        assert np.array_equal(sigs.view(np.uint64), z["sigs"].view(np.uint64)) and np.array_equal(cards, z["cards"])


def test_cmp_topk_csr_file(tmp_path):
    from dashing2_b200 import synth
    z = np.load(os.path.join(GOLD, "inputs", "sk600x64.npz"))
    stk = str(tmp_path / "sk600.ss")
    synth.write_stacked(stk, z["regs"], z["cards"], names=[f"s{i}" for i in range(600)])
    for K in (5, 32):
        out = str(tmp_path / f"top{K}.csr")
        run(["cmp", "--presketched", "--binary-output", "--topk", str(K), "--cmpout", out, stk])
        assert open(out, "rb").read() == open(expected(f"topk{K}_sk600.csr"), "rb").read()
        for nlsh in ("1", "3", "4", "5"):
            run(["cmp", "--presketched", "--binary-output", "--nLSH", nlsh, "--topk", str(K), "--cmpout", out, stk])
            assert open(out, "rb").read() == open(expected(f"topk{K}_nlsh{nlsh}_sk600.csr"), "rb").read()


def test_unsupported_options_fail_loudly(golden_inputs):
    names, paths = golden_inputs
    for argv in (["sketch", "-k31", "--full-setsketch", "-m", "2", paths[0]], ["sketch", "-k4000", paths[0]], ["sketch", "--protein", "-k15", "--parse-by-seq", paths[0]], ["sketch", "--parse-by-seq", paths[0], paths[1]],
                 ["wsketch", paths[0]]):
        r = subprocess.run([EXE] + argv, capture_output=True, text=True)
        assert r.returncode != 0 and r.stderr.strip()


def test_parse_by_seq_full_setsketch_on_a_read_set(tmp_path):
    """`sketch --parse-by-seq --full-setsketch -S4096` over 100 000 reads of 150 bp (failed with a queue overflow above
    ~10 000 reads): every element of such a record walks all 4096 registers.  A sample of records against the oracle."""
    import oracle_lib as O
    rng = np.random.default_rng(2026)
    n, L, S = 100_000, 150, 4096
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    reads = acgt[rng.integers(0, 4, size=(n, L), dtype=np.uint8)]
    fa = tmp_path / "reads.fa"
    with open(fa, "wb") as f:
        hdr = np.array([b">r%07d\n" % i for i in range(n)])
        body = np.concatenate([reads, np.full((n, 1), 10, dtype=np.uint8)], axis=1)
        for i in range(0, n, 10000):
            f.write(b"".join(h + b.tobytes() for h, b in zip(hdr[i:i + 10000], body[i:i + 10000])))
    out = str(tmp_path / "reads.stk")
    run(["sketch", "--parse-by-seq", "--full-setsketch", "-k31", f"-S{S}", "-o", out, str(fa)])
    n_out, s_out = (int(x) for x in np.fromfile(out, dtype=np.uint64, count=2))
    assert (n_out, s_out) == (n, S)
    sigs = np.memmap(out, dtype=np.float64, mode="r", offset=16 + 8 * n, shape=(n, S))
    sample = [0, 1, 4999, 65535, 65536, 65537, n - 1] + [int(x) for x in rng.integers(0, n, 25)]
    _, exp = O.sketch_records_byseq([reads[i].tobytes() for i in sample], "fss", S, 31, -1)
    for j, i in enumerate(sample):
        assert np.array_equal(np.asarray(sigs[i]).view(np.uint64), exp[j].view(np.uint64)), i


@pytest.mark.parametrize("case,argv", [("roll_opmh_k40_S128", ["-k40", "-S128"]), ("roll_opmh_k33_w50_S64_nocanon", ["-k33", "-w50", "-S64", "-C"]),
                                       ("roll_fss_k45_S64_seed3", ["-k45", "-S64", "--seed", "3", "--full-setsketch"]),
                                       ("opmh_k21_w30_S256_nocanon", ["-k21", "-w30", "-S256", "-C"])])
def test_element_stream_modes_stacked_file(case, argv, golden_inputs, tmp_path):
    """k > 32 (rolling hash) and -C with a window through the front-end: registers as the reference binary wrote them."""
    if case.startswith("roll"):
        files = ["g0.fa.gz", "g1.fa.gz", "dup.fa.gz", "adv.fa.gz", "reads.fq.gz"]
        paths = [os.path.join(GOLD, "inputs", f) for f in files]
    else:
        paths = golden_inputs[1]
    flist = tmp_path / "files.txt"; flist.write_text("\n".join(paths) + "\n")
    out = str(tmp_path / "out.stk")
    run(["sketch", "-p4", "-F", str(flist), "-o", out] + argv)
    cards, sigs = read_stacked(out)
    z = np.load(expected(case + ".npz"))
    assert np.array_equal(sigs.view(np.uint64), z["sigs"].view(np.uint64))
    if "fss" in case:
        np.testing.assert_allclose(cards, z["cards"], rtol=1e-12)
    else:
        assert np.array_equal(cards, z["cards"])


@pytest.mark.parametrize("case,argv", [("prot20_opmh_k7_S64", ["--protein", "-k7", "-S64"]), ("prot8_opmh_k12_S64", ["--protein8", "-k12", "-S64"]),
                                       ("prot14_opmh_k10_S64", ["--protein14", "-k10", "-S64"]), ("prot6_opmh_k20_S64", ["--protein6", "-k20", "-S64"]),
                                       ("prot20_opmh_k5_w12_S32", ["--protein", "-k5", "-w12", "-S32"]), ("prot20_fss_k7_S64", ["--protein", "-k7", "-S64", "--full-setsketch"])])
def test_protein_alphabets_stacked_files(case, argv, tmp_path):
    """--protein* through the front-end: --parse-by-seq sketches per record; per FILE the reference binary leaves the registers empty
    (reproduced).  Both stacked files as the reference binary wrote them."""
    import gzip
    fa = str(tmp_path / "prot.fa")
    open(fa, "wb").write(gzip.open(os.path.join(GOLD, "inputs", "prot.fa.gz"), "rb").read())
    z = np.load(expected(case + ".npz"))
    out = str(tmp_path / "byseq.stk")
    run(["sketch", "--parse-by-seq", "-p2", "-o", out] + argv + [fa])
    cards, sigs = read_stacked(out)
    assert np.array_equal(sigs.view(np.uint64), z["byseq_sigs"].view(np.uint64))
    if "fss" in case:
        np.testing.assert_allclose(cards, z["byseq_cards"], rtol=1e-12)
    else:
        assert np.array_equal(cards, z["byseq_cards"])
    out = str(tmp_path / "file.stk")
    run(["sketch", "-p2", "-o", out] + argv + [fa])
    cards, sigs = read_stacked(out)
    assert np.array_equal(sigs.view(np.uint64), z["sigs"].view(np.uint64)) and np.array_equal(cards, z["cards"])


def test_cmp_threshold_graph_and_topk_fastcmp_csr_files(tmp_path):
    """`cmp --similarity-threshold x` and `cmp --topk 8 --fastcmp N [--bbit-sigs]`: the CSR files of the reference binary, byte for byte."""
    from dashing2_b200 import synth
    z = np.load(os.path.join(GOLD, "inputs", "sk600x64.npz"))
    stk = str(tmp_path / "sk600.ss")
    synth.write_stacked(stk, z["regs"], z["cards"], names=[f"s{i}" for i in range(600)])
    out = str(tmp_path / "g.csr")
    for tag, argv in (("t0.5", ["--similarity-threshold", "0.5"]), ("t0.8", ["-T", "0.8"]), ("t0.3_containment", ["--similarity-threshold", "0.3", "--containment"])):
        run(["cmp", "--presketched", "--binary-output", "--cmpout", out, stk] + argv)
        assert open(out, "rb").read() == open(expected(f"nnthr_{tag}_sk600.csr"), "rb").read(), tag
    for tag, argv in (("fd1", ["--fastcmp", "1"]), ("fd2_bbit", ["--fastcmp", "2", "--bbit-sigs"])):
        run(["cmp", "--presketched", "--binary-output", "--topk", "8", "--cmpout", out, stk] + argv)
        assert open(out, "rb").read() == open(expected(f"topk8_{tag}_sk600.csr"), "rb").read(), tag


@pytest.mark.parametrize("case,argv", [("fs_opmh_k31_S128", ["-k31", "-S128"]), ("fs_opmh_k21_w30_S64", ["-k21", "-w30", "-S64"]),
                                       ("fs_fss_k31_S64", ["-k31", "-S64", "--full-setsketch"]), ("fs_opmh_k40_S64", ["-k40", "-S64"])])
def test_filterset_option(case, argv, tmp_path):
    """`sketch --filterset dup.fa`: registers as the reference binary wrote them; the combinations the reference binary itself cannot run
    (raw k-mer file, --multiset) are refused."""
    import gzip
    paths = []
    for f in ["g0.fa", "g1.fa", "dup.fa", "adv.fa"]:
        dst = str(tmp_path / f); open(dst, "wb").write(gzip.open(os.path.join(GOLD, "inputs", f + ".gz"), "rb").read()); paths.append(dst)
    flist = tmp_path / "files.txt"; flist.write_text("\n".join(paths) + "\n")
    out = str(tmp_path / "out.stk")
    run(["sketch", "-p4", "-F", str(flist), "-o", out, "--filterset", paths[2]] + argv)
    cards, sigs = read_stacked(out)
    z = np.load(expected(case + ".npz"))
    assert np.array_equal(sigs.view(np.uint64), z["sigs"].view(np.uint64))
    if "fss" in case:
        np.testing.assert_allclose(cards, z["cards"], rtol=1e-12)
    else:
        assert np.array_equal(cards, z["cards"])
    if case == "fs_opmh_k31_S128":
        for bad in (["--filterset", paths[2] + ":B"], ["--filterset", paths[2], "--multiset"]):
            r = subprocess.run([EXE, "sketch", "-F", str(flist), "-o", out, "-k31"] + bad, capture_output=True, text=True)
            assert r.returncode != 0 and r.stderr.strip()


@pytest.mark.parametrize("case,argv", [("contain_opmh_k31_S64", ["-k31", "-S64"]), ("contain_opmh_k21_w30_S32_seed5", ["-k21", "-w30", "-S32", "--seed", "5"]),
                                       ("contain_fss_k31_S64", ["-k31", "-S64", "--full-setsketch"])])
def test_contain_subcommand(case, argv, tmp_path):
    """`sketch --save-kmers` then `contain`: the database file and the coverage / mean-depth matrices as the reference binary wrote them
    (binary form; the text form for the case whose golden holds it)."""
    import gzip
    def fetch(f):
        dst = str(tmp_path / f); open(dst, "wb").write(gzip.open(os.path.join(GOLD, "inputs", f + ".gz"), "rb").read()); return dst
    refs = [fetch(f) for f in ["g0.fa", "g1.fa", "dup.fa", "adv.fa"]]
    queries = [fetch(f) for f in ["g0.fa", "rep.fa", "reads.fq", "adv.fa", "dup.fa"]]
    z = np.load(expected(case + ".npz"))
    db = str(tmp_path / "db.stk")
    run(["sketch", "-p1", "--save-kmers", "-o", db] + argv + refs)
    assert np.array_equal(np.fromfile(db + ".kmer64", dtype=np.uint32, count=4), z["hdr"])
    assert np.array_equal(np.fromfile(db + ".kmer64", dtype=np.uint64, offset=24).reshape(4, -1), z["ids"])
    out = str(tmp_path / "c.bin")
    qlist = tmp_path / "q.txt"; qlist.write_text("\n".join(queries[:2]) + "\n")
    run(["contain", "-b", "-o", out, "-F", str(qlist), db + ".kmer64"] + queries[2:])
    raw = open(out, "rb").read()
    assert tuple(np.frombuffer(raw, np.uint64, 2)) == (4, 5)
    mat = np.frombuffer(raw, np.float32, 40, 16).reshape(2, 5, 4)
    assert np.array_equal(mat[0].view(np.uint32), z["coverage"].view(np.uint32)) and np.array_equal(mat[1], z["depth"])
    if os.path.exists(expected(case + ".txt")):
        outt = str(tmp_path / "c.txt")
        run(["contain", "-o", outt, db + ".kmer64"] + queries)
        assert open(outt).read().replace(str(tmp_path) + "/", "") == open(expected(case + ".txt")).read()


def test_gpu_stats_json(golden_inputs, tmp_path):
    """--gpu-stats FILE: kernel launches, per-class device times and host phases of the run as JSON (an operator hook, not a reference option)."""
    import json
    names, paths = golden_inputs
    flist = tmp_path / "files.txt"; flist.write_text("\n".join(paths) + "\n")
    st = str(tmp_path / "stats.json")
    run(["sketch", "-F", str(flist), "-k31", "-w51", "-S512", "--full-setsketch", "-o", str(tmp_path / "o.stk"), "--cmpout", str(tmp_path / "o.f32"), "--binary-output", "--gpu-stats", st])
    d = json.load(open(st))
    dev = d["devices"][0]
    assert dev["kernel_launches"] > 5 and dev["kernel_ms"]["sketch_main"]["ms"] > 0 and dev["kernel_ms"]["sketch_main"]["timed_regions"] >= 1
    assert any(p["phase"].startswith("compare") for p in d["host_phases_ms"])
