"""Pins the oracle (oracle/d2_oracle.c) against outputs of the unmodified reference binary
(tests/golden/expected, produced by tests/golden/make_golden.py) and the two known answers the
reference's own tests hold for this path (bonsai/test/encoding.cpp:84,122)."""
import os

import numpy as np
import pytest

import oracle_lib as O
from conftest import expected, GOLD

SKETCH = {
    "opmh_k31_S1024": dict(mode="opmh", S=1024, k=31),
    "opmh_k31_w51_S512": dict(mode="opmh", S=512, k=31, w=51),
    "opmh_k21_S256_nocanon": dict(mode="opmh", S=256, k=21, canon=False),
    "opmh_k21_w30_S256_nocanon": dict(mode="opmh", S=256, k=21, w=30, canon=False),
    "opmh_k31_S256_seed17": dict(mode="opmh", S=256, k=31, seed=17),
    "opmh_k15_S64": dict(mode="opmh", S=64, k=15),
    "opmh_k31_S1000": dict(mode="opmh", S=1000, k=31),
    "fss_k31_S256": dict(mode="fss", S=256, k=31),
    "fss_k31_w51_S1024": dict(mode="fss", S=1024, k=31, w=51),
    "pmh_k31_S128": dict(mode="pmh", S=128, k=31),
    "bmh_k31_S128": dict(mode="bmh", S=128, k=31),
}


@pytest.mark.parametrize("case", sorted(SKETCH))
def test_sketch_registers_bit_exact(case, golden_inputs):
    names, paths = golden_inputs
    z = np.load(expected(case + ".npz"))
    for i, p in enumerate(paths):
        o = O.sketch_file(p, **SKETCH[case])
        assert np.array_equal(o["sig"].view(np.uint64), z["sigs"][i].view(np.uint64)), (case, names[i])
        if SKETCH[case]["mode"] in ("opmh", "pmh", "bmh"):
            assert o["card"] == z["cards"][i], (case, names[i])
        else:  # reference sums registers under `omp simd` (order is compiler-chosen): 1e-12 relative
            assert o["card"] == pytest.approx(z["cards"][i], rel=1e-12)


def test_save_kmers_ids(golden_inputs):
    names, paths = golden_inputs
    z = np.load(expected("opmh_k31_S256_savekmers.npz"))
    for i, p in enumerate(paths):
        o = O.sketch_file(p, mode="opmh", S=256, k=31)
        assert np.array_equal(o["ids"], z["ids"][i])


def test_densify_matches_reference(golden_inputs):
    names, paths = golden_inputs
    raw = np.load(expected("opmh_k15_S64.npz"))        # sketch only: not densified
    den = np.load(expected("opmh_k15_S64_densified.npz"))  # sketch --cmpout: densified in place
    changed = 0
    for i in range(len(paths)):
        got = O.densify(raw["sigs"][i])
        assert np.array_equal(got, den["sigs"][i])
        changed += int((raw["sigs"][i] != den["sigs"][i]).sum())
    assert changed > 0  # the fixture really exercises densification
    got = O.allpairs(den["sigs"], den["cards"], "symmetric", "similarity", k=15)
    assert np.array_equal(got.view(np.uint32), den["mat"].view(np.uint32))


CMP = {"sim_sym": ("symmetric", "similarity"), "sim_asym": ("asymmetric", "similarity"),
       "containment_sym": ("symmetric", "containment"), "symcontainment_sym": ("symmetric", "symmetric_containment"),
       "mash_sym": ("symmetric", "poisson_llr"), "isz_sym": ("symmetric", "intersection"),
       "usz_sym": ("symmetric", "union_size")}


@pytest.mark.parametrize("kind", sorted(CMP))
def test_compare_opmh_matrix(kind):
    z = np.load(expected("opmh_k31_S1024.npz"))
    sigs = np.stack([O.densify(s) for s in z["sigs"]])
    got = O.allpairs(sigs, z["cards"], CMP[kind][0], CMP[kind][1], k=31)
    exp = np.load(expected(f"cmp_opmh_k31_S1024_{kind}.npy"))
    assert np.array_equal(got.view(np.uint32), exp.view(np.uint32))


@pytest.mark.parametrize("suffix,cmp_kind", [(".ss", 0), (".bmh", 1)])
@pytest.mark.parametrize("kind", sorted(CMP))
def test_compare_presketched(kind, suffix, cmp_kind):
    f = expected(f"cmp_sk48{suffix}_{kind}.npy")
    if not os.path.exists(f):
        pytest.skip("no golden for this combination")
    z = np.load(os.path.join(os.path.dirname(expected("x")), "..", "inputs", "sk48x256.npz"))
    # `cmp --presketched` without -k: k defaults to 32 (src/sketch_main.cpp:70 nregperitem)
    got = O.allpairs(z["regs"], z["cards"], CMP[kind][0], CMP[kind][1], k=32, cmp_kind=cmp_kind)
    exp = np.load(f)
    assert np.array_equal(got.view(np.uint32), exp.view(np.uint32))


def test_known_answers_from_reference_tests():
    """bonsai/test/encoding.cpp:84 -- windowed minimizer count == len - w + 1 (canonical path);
    :122 pins 5356 distinct phiX 31-mers, whose fixture (phix.fa) is not redistributable here, so we
    check the structural identity it relies on: #k-mers emitted == len - k + 1 on ACGT-only input."""
    rng = np.random.default_rng(3)
    seq = b"ACGT"[0:0] + bytes(np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, 5386)])
    assert len(O.hash_stream(seq, 31)) == 5386 - 31 + 1
    for w in (40, 51, 100):
        assert len(O.hash_stream(seq, 31, w)) == 5386 - w + 1


def test_wang_inverse_roundtrip():
    L = O.lib()
    for x in (0, 1, 133348, 0xdeadbeefcafebabe, 2**64 - 1):
        assert L.d2o_wang64_inv(L.d2o_wang64(x)) == x


@pytest.mark.parametrize("K", [5, 32])
def test_topk_csr_matches_reference(K):
    """LSH candidate scan + bounded lists + refinement == `dashing2 cmp --presketched --topk K -p1` CSR."""
    z = np.load(os.path.join(os.path.dirname(expected("x")), "..", "inputs", "sk600x64.npz"))
    ip, ix, dv = O.read_csr(expected(f"topk{K}_sk600.csr"))
    gp, gi, gv = O.topk(z["regs"], z["cards"], K, "similarity", k=32)
    assert np.array_equal(gp, ip) and np.array_equal(gi, ix) and np.array_equal(gv.view(np.uint32), dv.view(np.uint32))


CMPC = {"sim_sym": ("symmetric", "similarity"), "sim_asym": ("asymmetric", "similarity"),
        "containment_sym": ("symmetric", "containment"), "symcontainment_sym": ("symmetric", "symmetric_containment"),
        "mash_sym": ("symmetric", "poisson_llr"), "isz_sym": ("symmetric", "intersection"), "usz_sym": ("symmetric", "union_size")}


@pytest.mark.parametrize("bbit", [False, True])
@pytest.mark.parametrize("fd", [1, 2, 4])
def test_compressed_compare_matches_reference(fd, bbit):
    """--fastcmp N [--bbit-sigs] on f64 registers: quantisation (a, b fitted from the data) + compressed compare,
    against the float32 matrices the reference binary wrote (tests/golden/make_golden_compressed.py)."""
    import json
    z = np.load(os.path.join(GOLD, "inputs", "sk48x256.npz"))
    creg, trunc, a, b = O.make_compressed(z["regs"], fd, bbit)
    assert trunc == (1 if bbit else 0)
    if not bbit:
        ab = json.load(open(expected("cmpc_fitted_ab.json")))[f"fd{fd}"]        # as printed by the reference with %0.20Lg / %0.24Lg
        import ctypes as C
        libc = C.CDLL(None); buf = C.create_string_buffer(128)
        libc.snprintf.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_longdouble, C.c_longdouble]
        libc.snprintf(buf, 128, b"%0.20Lg %0.24Lg", a, b)
        assert buf.value.decode().split() == ab
    for kind, (shape, measure) in CMPC.items():
        exp = np.load(expected(f"cmpc_sk48_fd{fd}_{'bbit' if bbit else 'ss'}_{kind}.npy"))
        got = O.allpairs_compressed(creg, z["cards"], shape, measure, fd, bbit, b, k=32)
        assert np.array_equal(got.view(np.uint32), exp.view(np.uint32)), (fd, bbit, kind)


BYSEQ = {
    "byseq_opmh_k31_S64": dict(mode="opmh", S=64, k=31),
    "byseq_opmh_k21_w30_S64": dict(mode="opmh", S=64, k=21, w=30),
    "byseq_opmh_k15_S16_nocanon": dict(mode="opmh", S=16, k=15, canon=False),
    "byseq_fss_k31_S64": dict(mode="fss", S=64, k=31),
    "byseq_fss_k21_w30_S32": dict(mode="fss", S=32, k=21, w=30),
    "byseq_bmh_k31_S32": dict(mode="bmh", S=32, k=31),
    "byseq_pmh_k31_S32": dict(mode="pmh", S=32, k=31),
}


@pytest.mark.parametrize("case", sorted(BYSEQ))
def test_parse_by_seq_matches_reference(case):
    """--parse-by-seq (src/fastxsketchbyseq.cpp): one sketch per record, exact distinct count below 10 * S, against the
    stacked files the reference binary wrote (tests/golden/make_golden_byseq.py)."""
    z = np.load(expected(case + ".npz"))
    recs = O.read_fastx(os.path.join(GOLD, "inputs", "byseq.fa.gz"))
    assert len(recs) == len(z["cards"]) == 16
    cards, sigs = O.sketch_records_byseq(recs, **BYSEQ[case])
    if "mat" in z.files and BYSEQ[case]["mode"] == "opmh":      # sketch --cmpout densifies the signatures in place before the file is closed
        sigs = np.stack([O.densify(s) for s in sigs])
    assert np.array_equal(sigs.view(np.uint64), z["sigs"].view(np.uint64))
    if BYSEQ[case]["mode"] == "fss":
        np.testing.assert_allclose(cards, z["cards"], rtol=1e-12)
        small = z["cards"] < 10 * BYSEQ[case]["S"]
        assert small.any() and np.array_equal(cards[small], z["cards"][small])     # the exact counts are exact
    else:
        assert np.array_equal(cards, z["cards"])
    if "mat" in z.files:
        got = O.allpairs(sigs, cards, "symmetric", "similarity", k=BYSEQ[case]["k"])
        assert np.array_equal(got.view(np.uint32), z["mat"].view(np.uint32))


MINCOUNT = {
    "mincount2_opmh_k31_S64": dict(S=64, k=31, count_threshold=2),
    "mincount3_opmh_k21_S128": dict(S=128, k=21, count_threshold=3),
    "mincount2_opmh_k21_w30_S64": dict(S=64, k=21, w=30, count_threshold=2),
}
MINCOUNT_FILES = ["rep.fa.gz", "dup.fa.gz", "g0.fa.gz", "adv.fa.gz"]


@pytest.mark.parametrize("case", sorted(MINCOUNT))
def test_opmh_count_threshold_matches_reference(case):
    """-m c (src/oph.h:188-205) per file, and the reference's --parse-by-seq quirk: its per-thread sketcher copies drop the
    mincount, so the one-permutation sketch of a record is the unfiltered one (tests/golden/make_golden_mincount.py)."""
    z = np.load(expected(case + ".npz"))
    kw = MINCOUNT[case]
    for i, f in enumerate(MINCOUNT_FILES):
        o = O.sketch_file(os.path.join(GOLD, "inputs", f), mode="opmh", **kw)
        assert np.array_equal(o["sig"].view(np.uint64), z["sigs"][i].view(np.uint64)) and o["card"] == z["cards"][i], (case, f)
    assert (z["sigs"] != 0).any()
    if "w" not in kw:
        assert (z["sigs"][2] == 0).all()                                # g0 has no repeated k-mer at all (every window counts when w > k)
    recs = O.read_fastx(os.path.join(GOLD, "inputs", "rep.fa.gz"))
    kw0 = {k: v for k, v in kw.items() if k != "count_threshold"}
    cards, sigs = O.sketch_records_byseq(recs, "opmh", **kw0)
    assert np.array_equal(sigs.view(np.uint64), z["byseq_sigs"].view(np.uint64)) and np.array_equal(cards, z["byseq_cards"])


PANEL = {"panel_opmh_k31_S1024_sim": ("opmh_k31_S1024", "similarity", True), "panel_opmh_k31_S1024_containment": ("opmh_k31_S1024", "containment", True),
         "panel_fss_k31_S256_mash": ("fss_k31_S256", "poisson_llr", False)}


@pytest.mark.parametrize("case", sorted(PANEL))
def test_panel_orientation_matches_reference(case):
    """-F refs -Q queries: rows = references, columns = queries (src/emitrect.cpp:229-246), against the matrix the reference binary
    wrote for 4 references x 5 queries (tests/golden/make_golden_panel.py)."""
    sk, measure, dens = PANEL[case]
    z = np.load(expected(sk + ".npz"))
    sigs = np.stack([O.densify(s) for s in z["sigs"]]) if dens else z["sigs"]
    got = O.allpairs(sigs, z["cards"], "panel", measure, k=31, nq=len(sigs) - 4)
    exp = np.load(expected(case + ".npy"))
    assert np.array_equal(got.view(np.uint32), exp.view(np.uint32))


COUNTSKETCH = {
    "cs5000_bmh_k31_S32": dict(mode="bmh", S=32, k=31, cssize=5000),
    "cs5000_pmh_k31_S32": dict(mode="pmh", S=32, k=31, cssize=5000),
    "cs300_pmh_k21_w30_S64": dict(mode="pmh", S=64, k=21, w=30, cssize=300),
    "cs100000_bmh_k31_S16": dict(mode="bmh", S=16, k=31, cssize=100000),
    "cs700_pmh_k31_S32_m3": dict(mode="pmh", S=32, k=31, cssize=700, count_threshold=3),
}
COUNTSKETCH_FILES = ["dup.fa.gz", "g0.fa.gz", "rep.fa.gz", "adv.fa.gz"]


@pytest.mark.parametrize("case", sorted(COUNTSKETCH))
def test_count_sketch_weighted_matches_reference(case):
    """--countsketch-size n (src/counter.h:68-77,131-137): signed float count sketch, elements (bucket index, |count|), threshold with >=."""
    z = np.load(expected(case + ".npz"))
    for i, f in enumerate(COUNTSKETCH_FILES):
        o = O.sketch_file(os.path.join(GOLD, "inputs", f), **COUNTSKETCH[case])
        assert np.array_equal(o["sig"].view(np.uint64), z["sigs"][i].view(np.uint64)) and o["card"] == z["cards"][i], (case, f)


WEIGHTED_IDS = {
    "ids_bmh_k31_S64": dict(mode="bmh", S=64, k=31),
    "ids_pmh_k31_S64": dict(mode="pmh", S=64, k=31),
    "ids_pmh_k21_w30_S32_seed5": dict(mode="pmh", S=32, k=21, w=30, seed=5),
    "ids_bmh_k15_S512": dict(mode="bmh", S=512, k=15),
}
WEIGHTED_IDS_FILES = ["dup.fa.gz", "g0.fa.gz", "rep.fa.gz", "adv.fa.gz", "reads.fq.gz"]


@pytest.mark.parametrize("case", sorted(WEIGHTED_IDS))
def test_weighted_save_kmers_ids_match_reference(case):
    """--save-kmers with --multiset / --prob: the id (maskfn'd k-mer) of the element that set each register (FILE.kmer64)."""
    z = np.load(expected(case + ".npz"))
    for i, f in enumerate(WEIGHTED_IDS_FILES):
        o = O.sketch_file(os.path.join(GOLD, "inputs", f), **WEIGHTED_IDS[case])
        assert np.array_equal(o["sig"].view(np.uint64), z["sigs"][i].view(np.uint64)) and o["card"] == z["cards"][i], (case, f)
        assert np.array_equal(o["ids"], z["ids"][i]), (case, f)


FSS_IDS = {"ids_fss_k31_S256": dict(S=256, k=31), "ids_fss_k21_w30_S64_seed5": dict(S=64, k=21, w=30, seed=5), "ids_fss_k15_S1024": dict(S=1024, k=15)}
FSS_IDS_FILES = ["dup.fa.gz", "g0.fa.gz", "g1.fa.gz", "adv.fa.gz", "reads.fq.gz"]


@pytest.mark.parametrize("case", sorted(FSS_IDS))
def test_fss_save_kmers_ids_match_reference(case):
    """--save-kmers --full-setsketch: the hashed k-mer that set each register (src/setsketch.h:400-404)."""
    z = np.load(expected(case + ".npz"))
    for i, f in enumerate(FSS_IDS_FILES):
        o = O.sketch_file(os.path.join(GOLD, "inputs", f), mode="fss", **FSS_IDS[case])
        assert np.array_equal(o["sig"].view(np.uint64), z["sigs"][i].view(np.uint64)), (case, f)
        assert np.array_equal(o["ids"], z["ids"][i]), (case, f)


@pytest.mark.parametrize("nlsh", [1, 3, 4, 5])
@pytest.mark.parametrize("K", [5, 32])
def test_topk_nlsh1_matches_reference(K, nlsh):
    """--nLSH 1: only the S one-register tables are built and scanned; --nLSH 3: 2S four-register tables on top, scanned first, three
    quarters of them keyed by XXH64 over wyhash-picked registers (src/cmp_core.cpp:757-770, src/ssi.h:355-392).  The GPU path has 1 and 2."""
    z = np.load(os.path.join(GOLD, "inputs", "sk600x64.npz"))
    ip, ix, dv = O.read_csr(expected(f"topk{K}_nlsh{nlsh}_sk600.csr"))
    gp, gi, gv = O.topk(z["regs"], z["cards"], K, "similarity", k=32, nlsh=nlsh)
    assert np.array_equal(gp, ip) and np.array_equal(gi, ix) and np.array_equal(gv.view(np.uint32), dv.view(np.uint32))


ROLLING = {
    "roll_opmh_k40_S128": dict(mode="opmh", S=128, k=40),
    "roll_opmh_k64_S64_nocanon": dict(mode="opmh", S=64, k=64, canon=False),
    "roll_opmh_k40_w60_S64": dict(mode="opmh", S=64, k=40, w=60),
    "roll_opmh_k33_w50_S64_nocanon": dict(mode="opmh", S=64, k=33, w=50, canon=False),
    "roll_fss_k45_S64_seed3": dict(mode="fss", S=64, k=45, seed=3),
}
ROLLING_FILES = ["g0.fa.gz", "g1.fa.gz", "dup.fa.gz", "adv.fa.gz", "reads.fq.gz"]


@pytest.mark.parametrize("case", sorted(ROLLING))
def test_rolling_hash_kmers_match_reference(case):
    """k > 32: the oracle's RollingHasher / CyclicHash stream (bonsai encoder.h:644-865) against sketches of the reference binary.
    Oracle only -- libd2gpu rejects k > 32; this pins the restatement the GPU path will be built against."""
    z = np.load(expected(case + ".npz"))
    for i, f in enumerate(ROLLING_FILES):
        o = O.sketch_file(os.path.join(GOLD, "inputs", f), **ROLLING[case])
        assert np.array_equal(o["sig"].view(np.uint64), z["sigs"][i].view(np.uint64)), (case, f)
        if ROLLING[case]["mode"] == "opmh":
            assert o["card"] == z["cards"][i], (case, f)


@pytest.mark.parametrize("tag,thr,measure", [("t0.5", 0.5, "similarity"), ("t0.8", 0.8, "similarity"), ("t0.3_containment", 0.3, "containment")])
def test_similarity_threshold_graph_matches_reference(tag, thr, measure):
    """--similarity-threshold x: candidate lists without a cap, ordered by hit count, refined with the 20-consecutive-failures rule
    (src/index_build.cpp:53-165 with topk = -1, src/refine.cpp:43-68).  Oracle only -- the GPU path for threshold graphs is not built."""
    z = np.load(os.path.join(GOLD, "inputs", "sk600x64.npz"))
    ip, ix, dv = O.read_csr(expected(f"nnthr_{tag}_sk600.csr"))
    gp, gi, gv = O.nn_threshold(z["regs"], z["cards"], thr, measure, k=32)
    assert np.array_equal(gp, ip) and np.array_equal(gi, ix) and np.array_equal(gv.view(np.uint32), dv.view(np.uint32))


PROTEIN = {
    "prot20_opmh_k7_S64": dict(mode="opmh", S=64, k=7, alphabet=20),
    "prot20_opmh_k14_S64": dict(mode="opmh", S=64, k=14, alphabet=20),
    "prot14_opmh_k10_S64": dict(mode="opmh", S=64, k=10, alphabet=14),
    "prot6_opmh_k20_S64": dict(mode="opmh", S=64, k=20, alphabet=6),
    "prot8_opmh_k12_S64": dict(mode="opmh", S=64, k=12, alphabet=8),
    "prot20_opmh_k5_w12_S32": dict(mode="opmh", S=32, k=5, w=12, alphabet=20),
    "prot20_fss_k7_S64": dict(mode="fss", S=64, k=7, alphabet=20),
}


@pytest.mark.parametrize("case", sorted(PROTEIN))
def test_protein_alphabets_match_reference(case):
    """--protein / --protein14 / --protein6 / --protein8 under --parse-by-seq: the oracle's protein k-mer stream (alphabet.h:107-120,
    rhtraits.h:52-62, encoder.h:241-306: `(min * mul) | code`, `% mul^k` or the k-bit mask of the 3-bit alphabet) against the reference
    binary.  Oracle only -- the GPU encode for these alphabets is not built.  Per FILE the reference binary sketches nothing for protein
    input (empty registers, recorded in the fixture), so only the per-record mode is pinned."""
    z = np.load(expected(case + ".npz"))
    recs = O.read_fastx(os.path.join(GOLD, "inputs", "prot.fa.gz"))
    cards, sigs = O.sketch_records_byseq(recs, canon=False, **PROTEIN[case])
    assert np.array_equal(sigs.view(np.uint64), z["byseq_sigs"].view(np.uint64))
    if PROTEIN[case]["mode"] == "opmh":
        assert np.array_equal(cards, z["byseq_cards"])
    else:
        np.testing.assert_allclose(cards, z["byseq_cards"], rtol=1e-12)
    assert (z["sigs"] == 0).all() or (z["sigs"] == np.finfo(np.float64).max).all()        # the per-file quirk: empty sketches


@pytest.mark.parametrize("case,kw", [("kmercounts_opmh_k31_S64", dict(S=64, k=31)), ("kmercounts_opmh_k21_w30_S128", dict(S=128, k=21, w=30))])
def test_opmh_kmercounts_match_reference(case, kw):
    """--save-kmercounts: multiplicity of each register's minimum (src/oph.h:206-209), float32 in FILE.kmercounts.f64.  Oracle only."""
    z = np.load(expected(case + ".npz"))
    for i, f in enumerate(["dup.fa.gz", "g0.fa.gz", "rep.fa.gz", "adv.fa.gz"]):
        o = O.sketch_file(os.path.join(GOLD, "inputs", f), mode="opmh", **kw)
        assert np.array_equal(o["counts"].astype(np.float32), z["counts"][i]), (case, f)


@pytest.mark.parametrize("case,kw", [("kmercounts_fss_k31_S64", dict(mode="fss", S=64, k=31)), ("kmercounts_bmh_k31_S32", dict(mode="bmh", S=32, k=31)),
                                     ("kmercounts_pmh_k31_S32", dict(mode="pmh", S=32, k=31))])
def test_owner_kmercounts_match_reference(case, kw):
    """--save-kmercounts for the other sketches: the count kept with a register is the multiplicity of the k-mer that owns it (its weight
    for BagMinHash / ProbMinHash), i.e. a function of the ids alone -- which is how a GPU pass would produce it.  Oracle only."""
    z = np.load(expected(case + ".npz"))
    for i, f in enumerate(["dup.fa.gz", "g0.fa.gz", "rep.fa.gz", "adv.fa.gz"]):
        path = os.path.join(GOLD, "inputs", f)
        o = O.sketch_file(path, **kw)
        hv = np.concatenate([O.hash_stream(r, kw["k"]) for r in O.read_fastx(path)])
        u, c = np.unique(hv, return_counts=True)
        mult = dict(zip(u.tolist(), c.tolist()))
        got = np.array([mult[int(x)] for x in o["ids"]], dtype=np.float32)
        assert np.array_equal(got, z["counts"][i]), (case, f)


@pytest.mark.parametrize("tag,fd,bbit", [("fd1", 1, False), ("fd2_bbit", 2, True)])
def test_topk_with_fastcmp_matches_reference(tag, fd, bbit):
    """--topk 8 --fastcmp N [--bbit-sigs]: LSH index over the f64 signatures, refinement through the compressed compare branch.
    Oracle only -- libd2gpu rejects the combination."""
    z = np.load(os.path.join(GOLD, "inputs", "sk600x64.npz"))
    creg, trunc, a, b = O.make_compressed(z["regs"], fd, bbit)
    ip, ix, dv = O.read_csr(expected(f"topk8_{tag}_sk600.csr"))
    gp, gi, gv = O.topk_compressed(z["regs"], creg, z["cards"], 8, fd, bbit, b, "similarity", k=32)
    assert np.array_equal(gp, ip) and np.array_equal(gi, ix) and np.array_equal(gv.view(np.uint32), dv.view(np.uint32))


FILTERSET = {
    "fs_opmh_k31_S128": dict(mode="opmh", S=128, k=31),
    "fs_opmh_k21_w30_S64": dict(mode="opmh", S=64, k=21, w=30),
    "fs_fss_k31_S64": dict(mode="fss", S=64, k=31),
    "fs_opmh_k40_S64": dict(mode="opmh", S=64, k=40),
}


@pytest.mark.parametrize("case", sorted(FILTERSET))
def test_filterset_matches_reference(case):
    """--filterset dup.fa (src/d2.cpp:45-98, src/fastxsketch.cpp:385-388): k-mers of the filter file never reach the sketch.  The reference
    binary crashes with --multiset and cannot open its raw k-mer file flavour (tests/golden/make_golden_filterset.py), so set sketches with
    a FASTX filter are what is pinned."""
    kw = FILTERSET[case]
    z = np.load(expected(case + ".npz"))
    fs = O.filterset_from_fastx(os.path.join(GOLD, "inputs", "dup.fa.gz"), kw["k"], kw.get("w", -1))
    for i, f in enumerate(["g0.fa.gz", "g1.fa.gz", "dup.fa.gz", "adv.fa.gz"]):
        o = O.sketch_file(os.path.join(GOLD, "inputs", f), filterset=fs, **kw)
        assert np.array_equal(o["sig"].view(np.uint64), z["sigs"][i].view(np.uint64)), (case, f)
        if kw["mode"] == "opmh":
            assert o["card"] == z["cards"][i], (case, f)
        else:
            np.testing.assert_allclose(o["card"], z["cards"][i], rtol=1e-12)


CONTAIN = {
    "contain_opmh_k31_S64": dict(k=31),
    "contain_opmh_k21_w30_S32_seed5": dict(k=21, w=30, seed=5),
    "contain_fss_k31_S64": dict(k=31),
}
CONTAIN_QUERIES = ["g0.fa.gz", "rep.fa.gz", "reads.fq.gz", "adv.fa.gz", "dup.fa.gz"]


@pytest.mark.parametrize("case", sorted(CONTAIN))
def test_contain_matches_reference(case):
    """`dashing2 contain` (src/contain_main.cpp:133-301): coverage and mean depth of a .kmer64 database's sampled k-mers in query streams."""
    z = np.load(expected(case + ".npz"))
    for qi, f in enumerate(CONTAIN_QUERIES):
        cov, depth = O.contain(z["ids"], O.read_fastx(os.path.join(GOLD, "inputs", f)), **CONTAIN[case])
        assert np.array_equal(cov.view(np.uint32), z["coverage"][qi].view(np.uint32)), (case, f)
        assert np.array_equal(depth, z["depth"][qi]), (case, f)
