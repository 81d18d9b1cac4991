"""GPU parity tests: the CUDA path (through the C ABI of libd2gpu.so) against the committed golden
vectors of the reference binary and against the oracle on seeded inputs.  Bit-exact for integer
registers, signatures and float32 matrices; 1e-12 relative only for the Full SetSketch cardinality
(the reference sums registers in compiler-chosen SIMD order)."""
import os

import numpy as np
import pytest

import oracle_lib as O
from conftest import expected, GOLD
from gpu_util import ctx, pack_batch, pack_files

pytestmark = pytest.mark.gpu

SKETCH = {
    "opmh_k31_S1024": dict(mode="opmh", S=1024, k=31),
    "opmh_k31_w51_S512": dict(mode="opmh", S=512, k=31, w=51),
    "opmh_k21_S256_nocanon": dict(mode="opmh", S=256, k=21, canon=False),
    "opmh_k31_S256_seed17": dict(mode="opmh", S=256, k=31, seed=17),
    "opmh_k15_S64": dict(mode="opmh", S=64, k=15),
    "opmh_k31_S1000": dict(mode="opmh", S=1000, k=31),
    "fss_k31_S256": dict(mode="fss", S=256, k=31),
    "fss_k31_w51_S1024": dict(mode="fss", S=1024, k=31, w=51),
    "pmh_k31_S128": dict(mode="pmh", S=128, k=31),
    "bmh_k31_S128": dict(mode="bmh", S=128, k=31),
}


def u64(a):
    return np.ascontiguousarray(a).view(np.uint64)


@pytest.mark.parametrize("case", sorted(SKETCH))
def test_sketch_matches_reference_golden(case, golden_inputs):
    names, paths = golden_inputs
    z = np.load(expected(case + ".npz"))
    c = ctx()
    seq, off, ent = pack_files(paths)
    r = c.sketch_batch(seq, off, ent, len(paths), c.params(**SKETCH[case]))
    for i, nm in enumerate(names):
        assert np.array_equal(u64(r["sig"][i]), u64(z["sigs"][i])), (case, nm)
    if SKETCH[case]["mode"] in ("opmh", "pmh", "bmh"):
        assert np.array_equal(u64(r["card"]), u64(z["cards"]))
    else:
        np.testing.assert_allclose(r["card"], z["cards"], rtol=1e-12)
    assert r["n_kmers"] == sum(max(0, len(rec) - SKETCH[case]["k"] + 1) for p in paths for rec in O.read_fastx(p))


def test_unsupported_modes_fail_loudly(golden_inputs):
    from dashing2_b200.capi import D2GError
    names, paths = golden_inputs
    c = ctx()
    seq, off, ent = pack_files(paths[:1])
    with pytest.raises(D2GError):
        c.sketch_batch(seq, off, ent, 1, c.params(mode="fss", S=64, k=21, count_threshold=2))
    with pytest.raises(D2GError):
        c.sketch_batch(seq, off, ent, 1, c.params(mode="opmh", S=64, k=4000))


def test_save_kmers_ids(golden_inputs):
    names, paths = golden_inputs
    z = np.load(expected("opmh_k31_S256_savekmers.npz"))
    c = ctx()
    seq, off, ent = pack_files(paths)
    r = c.sketch_batch(seq, off, ent, len(paths), c.params(mode="opmh", S=256, k=31), want_ids=True)
    assert np.array_equal(r["ids"], z["ids"])


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
def test_parse_by_seq_matches_reference_golden(case):
    """--parse-by-seq (src/fastxsketchbyseq.cpp:284-531): records are the entities of d2g_sketch_batch (empty records, records
    shorter than k / w, Ns, entity changes inside a tile); estimates below 10 * S are replaced by d2g_distinct_kmers."""
    kw = BYSEQ[case]; S = kw["S"]
    z = np.load(expected(case + ".npz"))
    recs = O.read_fastx(os.path.join(GOLD, "inputs", "byseq.fa.gz"))
    c = ctx()
    seq, off, ent = pack_batch([[r] for r in recs])
    p = c.params(**kw)
    r = c.sketch_batch(seq, off, ent, len(recs), p)
    sig, card = r["sig"], r["card"].copy()
    if "mat" in z.files and kw["mode"] == "opmh":
        sig = c.densify(sig)
    assert np.array_equal(u64(sig), u64(z["sigs"]))
    if kw["mode"] in ("opmh", "fss"):
        card[np.isnan(card)] = 0.
        small = np.flatnonzero(card < 10 * S)
        assert len(small) >= 6
        sseq, soff, sent = pack_batch([[recs[i]] for i in small])
        d = c.distinct_kmers(sseq, soff, sent, len(small), p)
        card[small] = d
        assert np.array_equal(card[small], z["cards"][small])
        # every record at once gives the same counts
        assert np.array_equal(c.distinct_kmers(seq, off, ent, len(recs), p)[small], d)
    if kw["mode"] == "fss":
        np.testing.assert_allclose(card, z["cards"], rtol=1e-12)
    else:
        assert np.array_equal(card, z["cards"])
    if "mat" in z.files:
        got = c.cmp_matrix(sig, card if kw["mode"] != "fss" else z["cards"], c.cmp_params(S, len(recs), "symmetric", "similarity", k=kw["k"]))
        assert np.array_equal(got.view(np.uint32), z["mat"].view(np.uint32))


MINCOUNT = {
    "mincount2_opmh_k31_S64": dict(S=64, k=31, count_threshold=2),
    "mincount3_opmh_k21_S128": dict(S=128, k=21, count_threshold=3),
    "mincount2_opmh_k21_w30_S64": dict(S=64, k=21, w=30, count_threshold=2),
}


@pytest.mark.parametrize("case", sorted(MINCOUNT))
def test_opmh_count_threshold_matches_reference_golden(case):
    """-m c with the one-permutation sketch (src/oph.h:188-205): register = minimum over the ids of the bucket seen >= c times."""
    z = np.load(expected(case + ".npz"))
    paths = [os.path.join(GOLD, "inputs", f) for f in ("rep.fa.gz", "dup.fa.gz", "g0.fa.gz", "adv.fa.gz")]
    c = ctx()
    seq, off, ent = pack_files(paths)
    r = c.sketch_batch(seq, off, ent, len(paths), c.params(mode="opmh", **MINCOUNT[case]))
    assert np.array_equal(u64(r["sig"]), u64(z["sigs"])) and np.array_equal(r["card"], z["cards"])


@pytest.mark.parametrize("S,k,w,thr", [(256, 31, -1, 2), (333, 17, 40, 3), (1024, 21, -1, 3), (64, 31, 51, 2)])
def test_opmh_count_threshold_matches_oracle_seeded(S, k, w, thr):
    from dashing2_b200 import synth
    files = []
    for g, s in synth.family_genomes(5, 40000, seed=300 + S, dup_frac=0.4):
        b = s.tobytes() if hasattr(s, "tobytes") else bytes(s)
        files.append([b[:25000], b[20000:] + b"N" + b[:3000], b"", b[100:100 + k - 1]])
    files.append([b""])                                   # an entity without a single k-mer
    c = ctx()
    seq, off, ent = pack_batch(files)
    r = c.sketch_batch(seq, off, ent, len(files), c.params(mode="opmh", S=S, k=k, w=w, count_threshold=thr))
    L = O.lib(); m = L.d2o_opmh_m(S)
    for e, recs in enumerate(files):
        hv = np.concatenate([O.hash_stream(x, k, w) for x in recs] + [np.empty(0, np.uint64)])
        regs = np.empty(m, dtype=np.uint64); counts = np.empty(m, dtype=np.float64)
        L.d2o_opmh_reset(regs, counts, m)
        L.d2o_opmh_update_mincount(regs, counts, m, hv, len(hv), float(thr))
        assert np.array_equal(r["regs_u64"][e], regs), e
        # order independence of the registers (the reference's candidate maps are order dependent only in the multiplicities)
        L.d2o_opmh_reset(regs, counts, m); hv2 = hv[::-1].copy()
        L.d2o_opmh_update_mincount(regs, counts, m, hv2, len(hv2), float(thr))
        assert np.array_equal(r["regs_u64"][e], regs), e
    assert (r["regs_u64"][:5] != np.uint64(2**64 - 1)).any()


WEIGHTED_IDS = {
    "ids_bmh_k31_S64": dict(mode="bmh", S=64, k=31),
    "ids_pmh_k31_S64": dict(mode="pmh", S=64, k=31),
    "ids_pmh_k21_w30_S32_seed5": dict(mode="pmh", S=32, k=21, w=30, seed=5),
    "ids_bmh_k15_S512": dict(mode="bmh", S=512, k=15),
}


@pytest.mark.parametrize("case", sorted(WEIGHTED_IDS))
def test_weighted_save_kmers_ids_match_reference_golden(case):
    """--save-kmers with --multiset / --prob: ids of the elements that set the registers (second pass over the final registers)."""
    z = np.load(expected(case + ".npz"))
    paths = [os.path.join(GOLD, "inputs", f) for f in ("dup.fa.gz", "g0.fa.gz", "rep.fa.gz", "adv.fa.gz", "reads.fq.gz")]
    c = ctx()
    seq, off, ent = pack_files(paths)
    r = c.sketch_batch(seq, off, ent, len(paths), c.params(**WEIGHTED_IDS[case]), want_ids=True)
    assert np.array_equal(u64(r["sig"]), u64(z["sigs"])) and np.array_equal(r["card"], z["cards"])
    assert np.array_equal(r["ids"], z["ids"])


@pytest.mark.parametrize("mode,S,k,w,cs", [("pmh", 1024, 31, -1, 0), ("bmh", 256, 21, 40, 0), ("pmh", 128, 31, -1, 2000), ("bmh", 2048, 31, -1, 0)])
def test_weighted_save_kmers_ids_match_oracle_seeded(mode, S, k, w, cs):
    from dashing2_b200 import synth
    files = []
    for g, s in synth.family_genomes(3, 30000, seed=700 + S, dup_frac=0.3):
        b = s.tobytes()
        files.append([b[:20000], b[15000:] + b"N" + b[:2000]])
    files.append([b"ACGT"])
    c = ctx()
    seq, off, ent = pack_batch(files)
    r = c.sketch_batch(seq, off, ent, len(files), c.params(mode=mode, S=S, k=k, w=w, cssize=cs), want_ids=True)
    for e, recs in enumerate(files[:-1]):
        hv = np.concatenate([O.hash_stream(x, k, w) for x in recs])
        o = O.weighted_sketch(hv, mode, S, 0, cs)
        assert np.array_equal(u64(r["sig"][e]), u64(o["sig"])) and np.array_equal(r["ids"][e], o["ids"]), e
    assert (r["ids"][-1] == 0).all()


FSS_IDS = {"ids_fss_k31_S256": dict(S=256, k=31), "ids_fss_k21_w30_S64_seed5": dict(S=64, k=21, w=30, seed=5), "ids_fss_k15_S1024": dict(S=1024, k=15)}


@pytest.mark.parametrize("case", sorted(FSS_IDS))
def test_fss_save_kmers_ids_match_reference_golden(case):
    """--save-kmers --full-setsketch: ids from the read-only second pass (FssIdsConsumer), including inputs far smaller than the sketch
    (reads.fq: every element walks all registers -> long-walk kernel)."""
    z = np.load(expected(case + ".npz"))
    paths = [os.path.join(GOLD, "inputs", f) for f in ("dup.fa.gz", "g0.fa.gz", "g1.fa.gz", "adv.fa.gz", "reads.fq.gz")]
    c = ctx()
    seq, off, ent = pack_files(paths)
    r = c.sketch_batch(seq, off, ent, len(paths), c.params(mode="fss", **FSS_IDS[case]), want_ids=True)
    assert np.array_equal(u64(r["sig"]), u64(z["sigs"]))
    assert np.array_equal(r["ids"], z["ids"])


@pytest.mark.parametrize("S,k,w,chunk", [(512, 31, -1, 0), (2048, 31, 51, 0), (128, 17, 40, 60000), (4096, 31, -1, 0)])
def test_fss_save_kmers_ids_match_oracle_seeded(S, k, w, chunk, monkeypatch, tmp_path):
    from dashing2_b200 import synth
    if chunk:
        monkeypatch.setenv("D2G_CHUNK_BYTES", str(chunk))      # several launches: ids land at their entity's offset
    files = []
    for g, s in synth.family_genomes(4, 60000, seed=900 + S):
        b = s.tobytes()
        files.append([b[:40000], b[40000:] + b"N" + b[:100]])
    files.append([b"ACGT"])
    c = ctx()
    seq, off, ent = pack_batch(files)
    r = c.sketch_batch(seq, off, ent, len(files), c.params(mode="fss", S=S, k=k, w=w), want_ids=True)
    for e, recs in enumerate(files[:-1]):
        f = tmp_path / f"e{e}.fa"
        f.write_bytes(b"".join(b">r\n" + x + b"\n" for x in recs))
        o = O.sketch_file(str(f), "fss", S, k, w)
        assert np.array_equal(u64(r["sig"][e]), u64(o["sig"])) and np.array_equal(r["ids"][e], o["ids"]), e
    assert (r["ids"][-1] == 0).all()


COUNTSKETCH = {
    "cs5000_bmh_k31_S32": dict(mode="bmh", S=32, k=31, cssize=5000),
    "cs5000_pmh_k31_S32": dict(mode="pmh", S=32, k=31, cssize=5000),
    "cs300_pmh_k21_w30_S64": dict(mode="pmh", S=64, k=21, w=30, cssize=300),
    "cs100000_bmh_k31_S16": dict(mode="bmh", S=16, k=31, cssize=100000),
    "cs700_pmh_k31_S32_m3": dict(mode="pmh", S=32, k=31, cssize=700, count_threshold=3),
}


@pytest.mark.parametrize("case", sorted(COUNTSKETCH))
def test_count_sketch_weighted_matches_reference_golden(case):
    """--countsketch-size n (src/counter.h:68-77,131-137) feeding BagMinHash / ProbMinHash: registers and total weights bit for bit."""
    z = np.load(expected(case + ".npz"))
    paths = [os.path.join(GOLD, "inputs", f) for f in ("dup.fa.gz", "g0.fa.gz", "rep.fa.gz", "adv.fa.gz")]
    c = ctx()
    seq, off, ent = pack_files(paths)
    r = c.sketch_batch(seq, off, ent, len(paths), c.params(**COUNTSKETCH[case]))
    assert np.array_equal(u64(r["sig"]), u64(z["sigs"])) and np.array_equal(r["card"], z["cards"])


@pytest.mark.parametrize("mode,S,k,w,cs,thr", [("pmh", 256, 31, -1, 4096, 0), ("bmh", 128, 21, 40, 1000, 2), ("pmh", 64, 17, -1, 37, 0), ("bmh", 64, 31, -1, 1 << 20, 0)])
def test_count_sketch_weighted_matches_oracle_seeded(mode, S, k, w, cs, thr):
    from dashing2_b200 import synth
    files = []
    for g, s in synth.family_genomes(4, 30000, seed=500 + S, dup_frac=0.3):
        b = s.tobytes()
        files.append([b[:20000], b[15000:] + b"N" + b[:2000], b""])
    files.append([b"ACGT"])                                # an entity without a single k-mer
    c = ctx()
    seq, off, ent = pack_batch(files)
    r = c.sketch_batch(seq, off, ent, len(files), c.params(mode=mode, S=S, k=k, w=w, cssize=cs, count_threshold=thr))
    for e, recs in enumerate(files):
        hv = np.concatenate([O.hash_stream(x, k, w) for x in recs] + [np.empty(0, np.uint64)])
        o = O.weighted_sketch(hv, mode, S, thr, cs)
        if len(hv) == 0:
            assert r["card"][e] == 0
            continue
        assert np.array_equal(u64(r["sig"][e]), u64(o["sig"])) and r["card"][e] == o["card"], e


@pytest.mark.parametrize("k,w,canon", [(31, -1, True), (21, 30, True), (15, -1, False), (11, 50, True), (32, -1, True)])
def test_distinct_kmers_matches_oracle_seeded(k, w, canon):
    """Exact distinct k-mers / minimizers per entity against the oracle's hashed stream, on reads with repeats, Ns, lower case,
    empty records, several records per entity and entities of very different sizes."""
    rng = np.random.default_rng(1000 + k + (w if w > 0 else 0))
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    ents = []
    for e in range(300):
        nrec = 1 if e % 7 else 3
        recs = []
        for _ in range(nrec):
            L = int(rng.choice([0, 5, k - 1, k, k + 3, 80, 150, 151, 1000, 20000 if e % 50 == 0 else 300]))
            a = acgt[rng.integers(0, 4 if e % 3 else 2, size=L)].copy()
            if L > 60 and e % 4 == 0:
                a[L // 2:L // 2 + 30] = a[:30]           # repeated content
            if L > 40 and e % 5 == 0:
                a[rng.integers(0, L, size=2)] = ord("N")
            b = a.tobytes()
            recs.append(b.lower() if e % 11 == 0 else b)
        ents.append(recs)
    c = ctx()
    seq, off, ent = pack_batch(ents)
    got = c.distinct_kmers(seq, off, ent, len(ents), c.params(mode="opmh", S=64, k=k, w=w, canon=canon))
    exp = np.array([len(np.unique(np.concatenate([O.hash_stream(r, k, w, canon) for r in recs] + [np.empty(0, np.uint64)]))) for recs in ents], dtype=np.uint64)
    assert np.array_equal(got, exp)
    # seed independence (maskfn is a bijection)
    assert np.array_equal(c.distinct_kmers(seq, off, ent, len(ents), c.params(mode="opmh", S=64, k=k, w=w, canon=canon, seed=9)), exp)


@pytest.mark.parametrize("mode,S,k,w", [("opmh", 64, 31, -1), ("opmh", 256, 21, 30), ("fss", 64, 31, -1), ("fss", 32, 15, 40)])
def test_many_small_entities_match_oracle(mode, S, k, w):
    """Thousands of read-sized entities in one batch (the --parse-by-seq shape): every record's registers against the oracle."""
    rng = np.random.default_rng(77 + S + k)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    recs = []
    for e in range(1500):
        L = int(rng.choice([0, 20, k, 100, 150, 250, 700]))
        a = acgt[rng.integers(0, 4, size=L)].copy()
        if L > 40 and e % 9 == 0:
            a[rng.integers(0, L)] = ord("N")
        recs.append(a.tobytes())
    c = ctx()
    seq, off, ent = pack_batch([[r] for r in recs])
    r = c.sketch_batch(seq, off, ent, len(recs), c.params(mode=mode, S=S, k=k, w=w))
    cards, sigs = O.sketch_records_byseq(recs, mode, S, k, w)
    assert np.array_equal(u64(r["sig"]), u64(sigs))


@pytest.mark.parametrize("mode,S,k,w", [("opmh", 1024, 31, -1), ("opmh", 4096, 31, 51), ("opmh", 333, 17, 40),
                                         ("fss", 512, 31, -1), ("fss", 2048, 31, 51), ("opmh", 8192, 32, -1),
                                         # window shapes: 2, 4 (< 8 keys: direct scan), exactly 8, 9, many groups of eight, > one thread's reach
                                         ("opmh", 256, 31, 32), ("opmh", 256, 31, 34), ("fss", 128, 31, 38), ("opmh", 512, 31, 39),
                                         ("opmh", 256, 31, 200), ("fss", 64, 15, 400), ("opmh", 128, 11, 1000)])
def test_sketch_matches_oracle_seeded(mode, S, k, w, tmp_path):
    """Seeded inputs larger than the goldens (multi-tile, multi-CTA, several entities per CTA span)."""
    from dashing2_b200 import synth
    rng = np.random.default_rng(1234 + S)
    files = []
    for g, s in synth.family_genomes(5, 150_000, seed=77 + S):
        b = bytearray(s.tobytes())
        for p in rng.integers(0, len(b), 12):       # sprinkle invalid bases / lowercase
            b[p] = ord("N")
        for p in rng.integers(0, len(b), 200):
            b[p] = b[p] | 0x20
        cut = sorted(rng.integers(1, len(b) - 1, 3))  # multi-record entity, one tiny record
        recs = [bytes(b[:cut[0]]), bytes(b[cut[0]:cut[1]]), b"ACGT", bytes(b[cut[1]:cut[2]]), b"", bytes(b[cut[2]:])]
        files.append(recs)
    files.append([b"ACGTAC"])          # entity with no k-mer at all
    files.append([])                    # entity with no record
    c = ctx()
    seq, off, ent = pack_batch(files)
    r = c.sketch_batch(seq, off, ent, len(files), c.params(mode=mode, S=S, k=k, w=w))
    L = O.lib()
    for e, recs in enumerate(files):
        hv = [O.hash_stream(x, k, w) for x in recs]
        hv = np.concatenate(hv) if hv else np.empty(0, dtype=np.uint64)
        if mode == "opmh":
            m = L.d2o_opmh_m(S)
            regs = np.empty(m, dtype=np.uint64); cnt = np.empty(m, dtype=np.float64)
            L.d2o_opmh_reset(regs, cnt, m); L.d2o_opmh_update(regs, cnt, m, hv, len(hv))
            assert np.array_equal(r["regs_u64"][e], regs), (mode, S, e)
            sig = np.empty(m, dtype=np.float64); L.d2o_opmh_sigs(regs, m, sig)
            assert np.array_equal(u64(r["sig"][e]), u64(sig[:S]))
            assert r["card"][e] == L.d2o_opmh_card(regs, m) or (np.isinf(r["card"][e]) and np.isinf(L.d2o_opmh_card(regs, m)))
        else:
            regs = np.empty(2 * S - 1, dtype=np.float64)
            L.d2o_css_reset(regs, S); L.d2o_css_update(regs, S, hv, len(hv), None)
            assert np.array_equal(u64(r["sig"][e]), u64(regs[:S])), (mode, S, e)
            np.testing.assert_allclose(r["card"][e], L.d2o_css_card(regs, S), rtol=1e-12)


@pytest.mark.parametrize("S,w", [(64, -1), (64, 40), (256, -1), (128, 51)])
def test_fss_guessed_bound_and_its_fallback(S, w, monkeypatch):
    """Full SetSketch takes its pruning bound from a guess (sequence length) that is verified afterwards.  Entities:
    random genomes (guess holds), a short tandem repeat (far fewer distinct elements than the length suggests: the guess
    fails and the entity is redone through the boot pass), inputs too small to guess for.  All bit-identical to the
    oracle, and to the library with guessing disabled."""
    from dashing2_b200 import synth
    k = 31
    rng = np.random.default_rng(S + 7)
    files = [[s.tobytes()] for _, s in synth.family_genomes(3, 150_000, seed=40 + S)]
    unit = bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), 173))
    files.append([unit * 900])                                       # ~156 kb of a 173-bp repeat
    files.append([unit * 400, files[0][0][:60_000]])                 # repeat + some random sequence
    files.append([files[1][0][:3000]])                               # small input
    c = ctx()
    seq, off, ent = pack_batch(files)
    p = c.params(mode="fss", S=S, k=k, w=w)
    r = c.sketch_batch(seq, off, ent, len(files), p)
    monkeypatch.setenv("D2G_FSS_NO_GUESS", "1")
    r0 = c.sketch_batch(seq, off, ent, len(files), p)
    assert np.array_equal(u64(r["sig"]), u64(r0["sig"])) and np.array_equal(u64(r["card"]), u64(r0["card"]))
    L = O.lib()
    for e, recs in enumerate(files):
        hv = np.concatenate([O.hash_stream(x, k, w) for x in recs])
        regs = np.empty(2 * S - 1, dtype=np.float64)
        L.d2o_css_reset(regs, S); L.d2o_css_update(regs, S, hv, len(hv), None)
        assert np.array_equal(u64(r["sig"][e]), u64(regs[:S])), (S, w, e)


@pytest.mark.parametrize("mode,S,k,w,thr", [("pmh", 256, 31, -1, 0), ("bmh", 256, 31, -1, 0), ("pmh", 1024, 21, 40, 0), ("bmh", 64, 21, 40, 0),
                                             ("pmh", 512, 31, -1, 1), ("bmh", 8192, 31, -1, 0), ("pmh", 8192, 31, -1, 0)])
def test_weighted_sketch_matches_oracle_seeded(mode, S, k, w, thr):
    """Counting sketches: duplicated content (counts > 1), windows (every window counts), --count-threshold,
    entities too small for the first bound guess, S up to the BASELINE config-3 value 8192."""
    from dashing2_b200 import synth
    files = []
    for g, s in synth.family_genomes(3, 120_000, seed=300 + S, dup_frac=0.3):
        b = s.tobytes()
        files.append([b[:70_000], b[70_000:], b"ACGTNNNN" + b[:500]])
    files.append([b"ACGT" * 20])        # tiny entity
    files.append([])
    c = ctx()
    seq, off, ent = pack_batch(files)
    r = c.sketch_batch(seq, off, ent, len(files), c.params(mode=mode, S=S, k=k, w=w, count_threshold=thr))
    for e, recs in enumerate(files):
        hv = [O.hash_stream(x, k, w) for x in recs]
        hv = np.concatenate(hv) if hv else np.empty(0, dtype=np.uint64)
        o = O.weighted_sketch(hv, mode, S, thr)
        assert np.array_equal(u64(r["sig"][e]), u64(o["sig"])), (mode, S, e, int((r["sig"][e] != o["sig"]).sum()))
        assert r["card"][e] == o["card"]


@pytest.mark.parametrize("mode,S,w", [("opmh", 512, -1), ("opmh", 256, 40), ("fss", 256, 51), ("fss", 128, -1)])
def test_sketch_chunked_upload_equals_single_upload(mode, S, w, monkeypatch):
    """d2g_sketch_batch uploads large batches in chunks of whole entities overlapped with the kernels; forcing tiny
    chunks (including empty entities between them) must give the same registers as one upload."""
    from dashing2_b200 import synth
    gen = [[s.tobytes()] for _, s in synth.family_genomes(7, 40_000, seed=11)]
    gen.insert(3, []); gen.insert(0, []); gen.append([])          # entities without records
    gen[5] = [gen[5][0][:15000], gen[5][0][15000:]]              # a multi-record entity
    seq, off, ent = pack_batch(gen)
    c = ctx()
    p = c.params(mode=mode, S=S, k=31, w=w)
    monkeypatch.setenv("D2G_CHUNK_BYTES", str(1 << 40)); a = c.sketch_batch(seq, off, ent, len(gen), p)
    monkeypatch.setenv("D2G_CHUNK_BYTES", "50000"); b = c.sketch_batch(seq, off, ent, len(gen), p)
    assert np.array_equal(u64(a["sig"]), u64(b["sig"])) and np.array_equal(u64(a["card"]), u64(b["card"]))
    if mode == "opmh":
        assert np.array_equal(a["regs_u64"], b["regs_u64"])


def test_sketch_merge_property_full_size():
    """Size-independent property at BASELINE scale (1 Mbp genomes, S=1024 / 4096): bucket minima are a
    min-monoid, so sketch(A ++ B) == min(sketch(A), sketch(B)) register-wise, and sketching the same
    bytes twice is idempotent."""
    from dashing2_b200 import synth
    c = ctx()
    gen = [s.tobytes() for _, s in synth.family_genomes(4, 1_000_000, seed=4242)]
    for kw in (dict(mode="opmh", S=1024, k=31), dict(mode="opmh", S=4096, k=31, w=51)):
        p = c.params(**kw)
        seq, off, ent = pack_batch([[gen[0]], [gen[1]], [gen[0], gen[1]], [gen[0], gen[0]]])
        r = c.sketch_batch(seq, off, ent, 4, p)["regs_u64"]
        assert np.array_equal(r[2], np.minimum(r[0], r[1]))
        assert np.array_equal(r[3], r[0])


def test_densify_matches_reference_golden():
    raw = np.load(expected("opmh_k15_S64.npz")); den = np.load(expected("opmh_k15_S64_densified.npz"))
    c = ctx()
    got = c.densify(raw["sigs"])
    assert np.array_equal(u64(got), u64(den["sigs"]))
    p = c.cmp_params(64, len(got), "symmetric", "similarity", k=15)
    assert np.array_equal(c.cmp_matrix(got, den["cards"], p).view(np.uint32), den["mat"].view(np.uint32))


@pytest.fixture(params=["f64", "codes", "codes-sorted"])
def cmp_path(request, monkeypatch):
    """Both comparison kernels: the f64 tile kernel and the 16-bit order-code kernel (cmp16_kernels.cuh), the latter with
    its codes from the shared-memory hash table where only != matters (default) and from the per-register sort."""
    monkeypatch.setenv("D2G_CMP_PATH", request.param.split("-")[0])
    if request.param == "codes-sorted":
        monkeypatch.setenv("D2G_C16_NO_HASH", "1")
    return request.param


CMP = {"sim_sym": ("symmetric", "similarity"), "sim_asym": ("asymmetric", "similarity"),
       "containment_sym": ("symmetric", "containment"), "symcontainment_sym": ("symmetric", "symmetric_containment"),
       "mash_sym": ("symmetric", "poisson_llr"), "isz_sym": ("symmetric", "intersection"),
       "usz_sym": ("symmetric", "union_size")}


@pytest.mark.parametrize("kind", sorted(CMP))
def test_compare_opmh_golden(kind, cmp_path):
    z = np.load(expected("opmh_k31_S1024.npz"))
    c = ctx()
    sigs = c.densify(z["sigs"])
    p = c.cmp_params(1024, len(sigs), CMP[kind][0], CMP[kind][1], k=31)
    got = c.cmp_matrix(sigs, z["cards"], p)
    exp = np.load(expected(f"cmp_opmh_k31_S1024_{kind}.npy"))
    assert np.array_equal(got.view(np.uint32), exp.view(np.uint32))


@pytest.mark.parametrize("suffix,cmp_kind", [(".ss", 0), (".bmh", 1)])
@pytest.mark.parametrize("kind", sorted(CMP))
def test_compare_presketched_golden(kind, suffix, cmp_kind, cmp_path):
    f = expected(f"cmp_sk48{suffix}_{kind}.npy")
    if not os.path.exists(f):
        pytest.skip("no golden for this combination")
    z = np.load(os.path.join(GOLD, "inputs", "sk48x256.npz"))
    c = ctx()
    p = c.cmp_params(256, 48, CMP[kind][0], CMP[kind][1], k=32, cmp_kind=cmp_kind)
    got = c.cmp_matrix(z["regs"], z["cards"], p)
    assert np.array_equal(got.view(np.uint32), np.load(f).view(np.uint32))


@pytest.mark.parametrize("bbit", [False, True])
@pytest.mark.parametrize("fd", [1, 2, 4])
def test_compressed_compare_golden(fd, bbit, cmp_path):
    """--fastcmp N [--bbit-sigs]: d2g_make_compressed + compressed compare against the reference binary's matrices."""
    z = np.load(os.path.join(GOLD, "inputs", "sk48x256.npz"))
    c = ctx()
    creg, kind, a, b = c.make_compressed(z["regs"], fd, bbit)
    oreg, trunc, oa, ob = O.make_compressed(z["regs"], fd, bbit)
    assert np.array_equal(creg, oreg) and kind == (3 if bbit else 2) and (bbit or (a.value == oa.value and b.value == ob.value))
    for name, (shape, measure) in CMP.items():
        p = c.cmp_params(256, 48, shape, measure, k=32, cmp_kind=kind, regbytes=fd, compressed_b=b)
        got = c.cmp_matrix(creg, z["cards"], p)
        exp = np.load(expected(f"cmpc_sk48_fd{fd}_{'bbit' if bbit else 'ss'}_{name}.npy"))
        assert np.array_equal(got.view(np.uint32), exp.view(np.uint32)), (fd, bbit, name)


@pytest.mark.parametrize("S,fd,bbit", [(1024, 1, False), (333, 2, False), (4096, 4, False), (512, 1, True), (1000, 2, True), (64, 4, True)])
def test_compressed_compare_matches_oracle_seeded(S, fd, bbit, cmp_path):
    from dashing2_b200 import synth
    n, nq = 150, 60
    regs, cards = synth.synthetic_sketches(n, S, seed=S + fd, n_families=6)
    regs *= 1e-3; regs[5, : S // 3] = 0.0; regs[9] = regs[8]
    cards = cards * (1 + np.arange(n) % 7) / 2.0
    c = ctx()
    creg, kind, a, b = c.make_compressed(regs, fd, bbit)
    oreg, trunc, oa, ob = O.make_compressed(regs, fd, bbit)
    assert np.array_equal(creg, oreg)
    for shape in ("symmetric", "panel"):
        for measure in ("similarity", "containment", "symmetric_containment", "poisson_llr", "intersection", "union_size"):
            p = c.cmp_params(S, n, shape, measure, k=21, cmp_kind=kind, nq=nq if shape == "panel" else 0, regbytes=fd, compressed_b=b)
            got = c.cmp_matrix(creg, cards, p)
            exp = O.allpairs_compressed(oreg, cards, shape, measure, fd, bbit, ob, k=21, nq=nq)
            assert np.array_equal(got.view(np.uint32), exp.view(np.uint32)), (S, fd, bbit, shape, measure)


@pytest.mark.parametrize("S", [1024, 1000, 4096, 77])
@pytest.mark.parametrize("shape", ["symmetric", "asymmetric", "panel"])
def test_compare_matches_oracle_seeded(S, shape, cmp_path):
    """Ragged sizes (n not a multiple of the 64-pair tile, S not a multiple of the 32-register chunk),
    all measures and both comparison kinds, against the oracle."""
    from dashing2_b200 import synth
    n, nq = 203, 71
    regs, cards = synth.synthetic_sketches(n, S, seed=S + 5, n_families=7)
    cards = cards * (1 + np.arange(n) % 9) / 3.0
    regs[3] = regs[4]                     # identical pair
    regs[10, : S // 2] = 0.0              # zeros
    c = ctx()
    for measure in ("similarity", "containment", "symmetric_containment", "poisson_llr", "intersection", "union_size"):
        for cmp_kind in (0, 1):
            p = c.cmp_params(S, n, shape, measure, k=31, cmp_kind=cmp_kind, nq=nq if shape == "panel" else 0)
            got = c.cmp_matrix(regs, cards, p)
            exp = O.allpairs(regs, cards, shape, measure, k=31, cmp_kind=cmp_kind, nq=nq)
            assert np.array_equal(got.view(np.uint32), exp.view(np.uint32)), (S, shape, measure, cmp_kind)


def test_compare_counts_and_stream_blocks(cmp_path):
    from dashing2_b200 import synth
    regs, cards = synth.synthetic_sketches(150, 512, seed=9, n_families=5)
    c = ctx()
    c0, c1 = c.cmp_counts(regs[:100], regs[100:])
    a, b = regs[:100, None, :], regs[None, 100:, :]
    assert np.array_equal(c0, (a > b).sum(-1)) and np.array_equal(c1, (a < b).sum(-1))
    # streamed row ranges concatenate to the full matrix
    p = c.cmp_params(512, 150, "symmetric")
    full = c.cmp_matrix(regs, cards, p)
    parts = []
    for r0, r1 in ((0, 10), (10, 99), (99, 150)):
        c.cmp_stream(regs, cards, p, r0, r1, lambda blk, fr, nr: parts.append(blk.copy()))
    assert np.array_equal(np.concatenate(parts), full)
    assert np.array_equal(np.concatenate([c.cmp_rows(regs, cards, p, a, b) for a, b in ((0, 70), (70, 71), (71, 150))]), full)


def test_compare_full_size_properties(cmp_path):
    """BASELINE-scale property checks (S=4096, n=1500): the condensed symmetric matrix equals the upper
    triangle of the asymmetric one; the asymmetric similarity matrix is symmetric with unit diagonal."""
    from dashing2_b200 import synth
    n, S = 1500, 4096
    regs, cards = synth.synthetic_sketches(n, S, seed=31, n_families=40)
    c = ctx()
    sym = c.cmp_matrix(regs, cards, c.cmp_params(S, n, "symmetric"))
    full = c.cmp_matrix(regs, cards, c.cmp_params(S, n, "asymmetric")).reshape(n, n)
    iu = np.triu_indices(n, 1)
    assert np.array_equal(sym, full[iu])
    assert np.array_equal(full, full.T) and np.all(np.diag(full) == 1.0)


@pytest.mark.parametrize("shape", ["symmetric", "asymmetric", "panel"])
@pytest.mark.parametrize("maxjob,global_ranks", [(256, True), (384, True), (384, False)])
def test_compare_codes_blocked_jobs(shape, maxjob, global_ranks, monkeypatch):
    """Order-code path with the per-job sketch limit forced down so one call is split into many block-pair
    jobs (diagonal, off-diagonal, ragged last blocks), streamed in row ranges.  Jobs take their codes from
    global ranks computed once per call, or (global_ranks False) sort their own sketches."""
    from dashing2_b200 import synth
    monkeypatch.setenv("D2G_CMP_PATH", "codes")
    monkeypatch.setenv("D2G_C16_MAXJOB", str(maxjob))
    if not global_ranks:
        monkeypatch.setenv("D2G_C16_NO_GLOBAL", "1")
    n, nq, S = 1000, 390, 130
    regs, cards = synth.synthetic_sketches(n, S, seed=77, n_families=9)
    c = ctx()
    for cmp_kind, measure in ((0, "containment"), (1, "similarity")):
        p = c.cmp_params(S, n, shape, measure, k=31, cmp_kind=cmp_kind, nq=nq if shape == "panel" else 0)
        exp = O.allpairs(regs, cards, shape, measure, k=31, cmp_kind=cmp_kind, nq=nq)
        assert np.array_equal(c.cmp_matrix(regs, cards, p).view(np.uint32), exp.view(np.uint32)), (shape, cmp_kind)
        nrows = n - nq if shape == "panel" else n
        parts = []
        for r0, r1 in ((0, 131), (131, 500), (500, nrows)):
            if r0 < r1:
                c.cmp_stream(regs, cards, p, r0, min(r1, nrows), lambda blk, fr, nr: parts.append(blk.copy()))
        assert np.array_equal(np.concatenate(parts).view(np.uint32), exp.view(np.uint32)), (shape, cmp_kind, "stream")


def test_compare_codes_special_values(monkeypatch):
    """Signed zeros, infinities, subnormals, negative registers and ties rank exactly like the doubles compare;
    NaN registers (unordered) make the job fall back to the f64 kernel on the device."""
    from dashing2_b200 import synth
    n, S = 300, 96
    rng = np.random.default_rng(5)
    pool = np.array([0.0, -0.0, np.inf, -np.inf, 5e-324, -5e-324, 1.0, -1.0, 2.5, 1e308, np.nextafter(1.0, 2.0)])
    regs = pool[rng.integers(0, len(pool), size=(n, S))]
    cards = rng.random(n) * 1e6 + 1
    c = ctx()
    for with_nan in (False, True):
        if with_nan:
            regs = regs.copy(); regs[7, 3] = np.nan; regs[250, 95] = np.nan
        for cmp_kind in (0, 1):
            p = c.cmp_params(S, n, "symmetric", "similarity", k=31, cmp_kind=cmp_kind)
            monkeypatch.setenv("D2G_CMP_PATH", "f64"); a = c.cmp_matrix(regs, cards, p)
            monkeypatch.setenv("D2G_CMP_PATH", "codes"); b = c.cmp_matrix(regs, cards, p)
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), (with_nan, cmp_kind)
            if not with_nan:
                exp = O.allpairs(regs, cards, "symmetric", "similarity", k=31, cmp_kind=cmp_kind)
                assert np.array_equal(b.view(np.uint32), exp.view(np.uint32)), cmp_kind


def test_compare_codes_more_sketches_than_ranks(monkeypatch):
    """70 000 column sketches (> 63 487 half codes): the panel is split into column blocks."""
    from dashing2_b200 import synth
    monkeypatch.setenv("D2G_CMP_PATH", "codes")
    nf, nq, S = 150, 70000, 64
    regs, cards = synth.synthetic_sketches(nf + nq, S, seed=12, n_families=50)
    c = ctx()
    p = c.cmp_params(S, nf + nq, "panel", "similarity", k=31, nq=nq)
    got = c.cmp_matrix(regs, cards, p)
    monkeypatch.setenv("D2G_CMP_PATH", "f64")
    assert np.array_equal(got.view(np.uint32), c.cmp_matrix(regs, cards, p).view(np.uint32))
    a, b = regs[:nf, None, :], regs[None, nf:nf + 2000, :]
    eq = S - (a > b).sum(-1) - (a < b).sum(-1)
    assert np.array_equal(got.reshape(nf, nq)[:, :2000], (eq / S).astype(np.float32))


@pytest.mark.parametrize("K", [5, 32])
def test_topk_matches_reference_golden(K):
    z = np.load(os.path.join(GOLD, "inputs", "sk600x64.npz"))
    ip, ix, dv = O.read_csr(expected(f"topk{K}_sk600.csr"))
    gp, gi, gv = ctx().lsh_topk(z["regs"], z["cards"], K, "similarity", k=32)
    assert np.array_equal(gp, ip) and np.array_equal(gi, ix) and np.array_equal(gv.view(np.uint32), dv.view(np.uint32))


@pytest.mark.parametrize("nlsh", [1, 3, 4, 5])
@pytest.mark.parametrize("K", [5, 32])
def test_topk_nlsh1_matches_reference_golden(K, nlsh):
    """--nLSH 1: the index holds only the S one-register tables; --nLSH 3: 2S four-register tables on top (hashmem256 / XXH64 keys),
    scanned first; --nLSH 4 / 5: 8S/6 six-register and S eight-register tables keyed by XXH3_64bits, or XXH64 over 6 / 16 wyhash-picked
    registers beyond the last whole group (src/cmp_core.cpp:757-770, src/ssi.h:345-392)."""
    z = np.load(os.path.join(GOLD, "inputs", "sk600x64.npz"))
    ip, ix, dv = O.read_csr(expected(f"topk{K}_nlsh{nlsh}_sk600.csr"))
    gp, gi, gv = ctx().lsh_topk(z["regs"], z["cards"], K, "similarity", k=32, nlsh=nlsh)
    assert np.array_equal(gp, ip) and np.array_equal(gi, ix) and np.array_equal(gv.view(np.uint32), dv.view(np.uint32))


def test_topk_nlsh1_matches_oracle_seeded_and_nlsh3_fails():
    from dashing2_b200 import synth
    from dashing2_b200.capi import D2GError
    regs, cards = synth.synthetic_sketches(2500, 128, seed=77, n_families=40)
    for nlsh in (1, 3, 4, 6, 9):
        ip, ix, dv = O.topk(regs, cards, 10, "similarity", k=31, nlsh=nlsh)
        gp, gi, gv = ctx().lsh_topk(regs, cards, 10, "similarity", k=31, nlsh=nlsh)
        assert np.array_equal(gp, ip) and np.array_equal(gi, ix) and np.array_equal(gv.view(np.uint32), dv.view(np.uint32)), nlsh
    with pytest.raises(D2GError):
        ctx().lsh_topk(regs[:100], cards[:100], 5, nlsh=10)


def test_topk_row_ranges_concatenate_to_the_graph():
    """d2g_lsh_topk_rows: the lists of a range of sketches (how several GPUs share one graph) concatenate to the golden CSR."""
    z = np.load(os.path.join(GOLD, "inputs", "sk600x64.npz"))
    ip, ix, dv = O.read_csr(expected("topk32_sk600.csr"))
    c = ctx()
    ptr = [np.zeros(1, dtype=np.uint64)]; idx = []; val = []; base = np.uint64(0)
    for r0, r1 in ((0, 100), (100, 101), (101, 101), (101, 600)):
        gp, gi, gv = c.lsh_topk(z["regs"], z["cards"], 32, "similarity", k=32, rows=(r0, r1))
        assert gp[0] == 0 and len(gp) == r1 - r0 + 1
        ptr.append(gp[1:] + base); base += gp[-1]; idx.append(gi); val.append(gv)
    assert np.array_equal(np.concatenate(ptr), ip) and np.array_equal(np.concatenate(idx), ix)
    assert np.array_equal(np.concatenate(val).view(np.uint32), dv.view(np.uint32))


@pytest.mark.parametrize("n,S,K,measure,cmp_kind", [(3000, 128, 10, "similarity", 0), (1500, 256, 32, "containment", 0),
                                                     (2000, 64, 7, "poisson_llr", 0), (1200, 128, 16, "similarity", 1),
                                                     (700, 1024, 32, "symmetric_containment", 0)])
def test_topk_matches_oracle_seeded(n, S, K, measure, cmp_kind):
    from dashing2_b200 import synth
    regs, cards = synth.synthetic_sketches(n, S, seed=n + S, n_families=max(2, n // 40), p_lo=0.02, p_hi=0.7)
    cards = cards * (1 + np.arange(n) % 3)
    ep, ei, ev = O.topk(regs, cards, K, measure, k=31, cmp_kind=cmp_kind)
    gp, gi, gv = ctx().lsh_topk(regs, cards, K, measure, k=31, cmp_kind=cmp_kind)
    assert np.array_equal(gp, ep) and np.array_equal(gi, ei) and np.array_equal(gv.view(np.uint32), ev.view(np.uint32))


@pytest.mark.parametrize("n,S,K", [(1300, 64, 500), (1400, 64, 1100)])
def test_topk_large_k_matches_oracle(n, S, K):
    """Candidate lists longer than 1536 entries (topk >= 440) need more dynamic shared memory than the default limit in the
    query kernel, and lists above 512 take the per-thread replay path; both against the oracle."""
    from dashing2_b200 import synth
    regs, cards = synth.synthetic_sketches(n, S, seed=n + K, n_families=3, p_lo=0.02, p_hi=0.5)
    ep, ei, ev = O.topk(regs, cards, K, "similarity", k=31)
    gp, gi, gv = ctx().lsh_topk(regs, cards, K, "similarity", k=31)
    assert np.array_equal(gp, ep) and np.array_equal(gi, ei) and np.array_equal(gv.view(np.uint32), ev.view(np.uint32))


@pytest.mark.parametrize("mode,S,w", [("opmh", 1024, -1), ("fss", 512, 51), ("opmh", 256, 40)])
def test_packed_entry_points_equal_ascii_entry_point(mode, S, w, golden_inputs):
    """d2g_pack_sequences + d2g_sketch_batch_packed (with and without the mask) and the device packer behind
    d2g_sketch_batch_dev give the registers of d2g_sketch_batch on the golden FASTA fixtures (N runs, lower case, short records)."""
    import torch
    from dashing2_b200 import capi
    names, paths = golden_inputs
    recs = [O.read_fastx(p) for p in paths]
    seq, off, ent = pack_batch(recs)
    c = ctx()
    p = c.params(mode=mode, S=S, k=31, w=w)
    a = c.sketch_batch(seq, off, ent, len(recs), p)
    codes, mask, nz = capi.pack_sequences([r for rr in recs for r in rr])
    assert nz > 0                                            # adv.fa holds N runs
    b = c.sketch_batch_packed(codes, mask, off, ent, len(recs), p)
    assert np.array_equal(u64(a["sig"]), u64(b["sig"])) and np.array_equal(u64(a["card"]), u64(b["card"]))
    # a batch without invalid bases may drop the mask
    clean = [np.frombuffer(r, dtype=np.uint8) for rr in recs[:3] for r in rr]
    clean = [x[np.isin(x & 0xDF, np.frombuffer(b"ACGT", dtype=np.uint8))].tobytes() for x in clean]
    seq2, off2, ent2 = pack_batch([[x] for x in clean])
    codes2, mask2, nz2 = capi.pack_sequences(clean)
    assert nz2 == 0
    a2 = c.sketch_batch(seq2, off2, ent2, len(clean), p)
    b2 = c.sketch_batch_packed(codes2, None, off2, ent2, len(clean), p)
    assert np.array_equal(u64(a2["sig"]), u64(b2["sig"]))
    # device packer == host packer, word for word
    dev = torch.device("cuda", 0)
    t_seq = torch.from_numpy(np.concatenate([seq, np.zeros(64, np.uint8)])).to(dev)
    nw = len(codes)
    t_codes = torch.empty(nw, dtype=torch.int64, device=dev); t_mask = torch.empty(nw, dtype=torch.int32, device=dev)
    c.pack_dev(t_seq.data_ptr(), len(seq), t_codes.data_ptr(), t_mask.data_ptr())
    c.sync()
    assert np.array_equal(t_codes.cpu().numpy().view(np.uint64), codes)
    assert np.array_equal(t_mask.cpu().numpy().view(np.uint32), mask)


# ---- the 32-bit-key windowed kernel (csrc/sketch_fast.cuh) against the exact kernel and the oracle ------------------------
def _fast_inputs(seed, n_long=3, length=260_000):
    """Entities that exercise every path of the fast windowed kernel: multi-tile random genomes with N runs / lower case /
    several records, homopolymer and short-period repeats (every window ties with its predecessor), a record exactly one
    window long, records shorter than the window, an empty entity."""
    from dashing2_b200 import synth
    rng = np.random.default_rng(seed)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    files = []
    for g, s in synth.family_genomes(n_long, length, seed=seed):
        b = bytearray(s.tobytes())
        for p in rng.integers(0, len(b), 10):
            b[p] = ord("N")
        p = int(rng.integers(1000, len(b) - 5000)); b[p:p + 700] = b"N" * 700            # an N run longer than a window
        for p in rng.integers(0, len(b), 100):
            b[p] |= 0x20
        cut = sorted(int(x) for x in rng.integers(1, len(b) - 1, 2))
        files.append([bytes(b[:cut[0]]), bytes(b[cut[0]:cut[1]]), bytes(b[cut[1]:])])
    files.append([b"A" * 9000 + acgt[rng.integers(0, 4, 5000)].tobytes() + b"AC" * 4000 + b"ACG" * 3000 + b"T" * 2100])
    unit = acgt[rng.integers(0, 4, 13)].tobytes()
    files.append([unit * 700, acgt[rng.integers(0, 4, 3000)].tobytes()])
    files.append([acgt[rng.integers(0, 4, 51)].tobytes(), acgt[rng.integers(0, 4, 50)].tobytes(), acgt[rng.integers(0, 4, 52)].tobytes()])
    files.append([])
    return files


def _oracle_regs(mode, S, k, w, recs):
    L = O.lib()
    hv = [O.hash_stream(x, k, w) for x in recs]
    hv = np.concatenate(hv) if hv else np.empty(0, dtype=np.uint64)
    if mode == "opmh":
        m = L.d2o_opmh_m(S)
        regs = np.empty(m, dtype=np.uint64); cnt = np.empty(m, dtype=np.float64)
        L.d2o_opmh_reset(regs, cnt, m); L.d2o_opmh_update(regs, cnt, m, hv, len(hv))
        return regs
    regs = np.empty(2 * S - 1, dtype=np.float64)
    L.d2o_css_reset(regs, S); L.d2o_css_update(regs, S, hv, len(hv), None)
    return regs[:S]


@pytest.mark.parametrize("mode,S,k,w", [("opmh", 1024, 31, 51), ("fss", 1024, 31, 51), ("opmh", 256, 21, 40), ("fss", 256, 31, 33),
                                         ("opmh", 512, 31, 93), ("opmh", 128, 17, 24), ("fss", 64, 32, 70), ("opmh", 64, 31, 32)])
def test_fast_windowed_kernel_matches_oracle_and_exact_kernel(mode, S, k, w, monkeypatch):
    files = _fast_inputs(900 + S + w)
    c = ctx()
    seq, off, ent = pack_batch(files)
    p = c.params(mode=mode, S=S, k=k, w=w)
    fast = c.sketch_batch(seq, off, ent, len(files), p)
    key = "regs_u64" if mode == "opmh" else "sig"
    for e, recs in enumerate(files):
        assert np.array_equal(u64(fast[key][e]), u64(_oracle_regs(mode, S, k, w, recs))), (mode, S, k, w, e)
    monkeypatch.setenv("D2G_NO_FAST", "1")
    exact = c.sketch_batch(seq, off, ent, len(files), p)
    assert np.array_equal(u64(fast[key]), u64(exact[key])) and np.array_equal(u64(fast["card"]), u64(exact["card"]))


@pytest.mark.parametrize("keymask", ["0xFFF00000", "0xFF000000", "0xC0000000", "0"])
@pytest.mark.parametrize("mode,S,w", [("opmh", 512, 51), ("fss", 256, 51), ("opmh", 256, 38)])
def test_fast_windowed_kernel_key_ties_are_resolved_exactly(mode, S, w, keymask, monkeypatch):
    """With only a few bits of the window key taking part, equal keys of different k-mers are everywhere: the scan has to pick
    the minimizer by the full score and the redo list has to catch minimizer changes the truncated keys cannot see; with mask 0
    every window is an event, the lists overflow and every tile is recomputed.  Registers must not change."""
    files = _fast_inputs(17 + S + w, n_long=2, length=120_000)
    c = ctx()
    seq, off, ent = pack_batch(files)
    p = c.params(mode=mode, S=S, k=31, w=w)
    ref = c.sketch_batch(seq, off, ent, len(files), p)
    monkeypatch.setenv("D2G_FAST_KEYMASK", keymask)
    got = c.sketch_batch(seq, off, ent, len(files), p)
    key = "regs_u64" if mode == "opmh" else "sig"
    assert np.array_equal(u64(got[key]), u64(ref[key]))
    for e, recs in enumerate(files[:2]):
        assert np.array_equal(u64(got[key][e]), u64(_oracle_regs(mode, S, 31, w, recs)))


def _frev64_inv(s):
    M = (1 << 64) - 1
    imul = pow(0x9a98567ed20c127d | 1, -1, 1 << 64)
    s ^= 0x691a9d706391077a
    s = ((s >> 31) | (s << 33)) & M
    s = (s * imul) & M
    return s ^ 0x533f8c2151b20f97


def _kmer_str(x, k):
    return "".join("ACGT"[(x >> (2 * (k - 1 - i))) & 3] for i in range(k))


def _revcomp_int(x, k):
    r = 0
    for i in range(k):
        r = (r << 2) | (3 - ((x >> (2 * i)) & 3))
    return r


def test_fast_windowed_kernel_real_32bit_key_tie():
    """Two different canonical 31-mers whose scores agree in the high 32 bits (and are smaller than everything around them), planted
    a few positions apart in random sequence in both orders: the window minimum changes from one to the other without the 32-bit
    minimum changing.  OPMH registers against the oracle."""
    k, w, S = 31, 51, 256
    rng = np.random.default_rng(5)
    found = []
    hi = 3                                            # a tiny score: the minimizer of every window that holds it
    lo = 1
    while len(found) < 2:
        x = _frev64_inv((hi << 32) | lo)
        lo += 1
        if x >> 62 or x > _revcomp_int(x, k):
            continue
        s = _kmer_str(x, k)
        if any(s[i:] == s[:-i] for i in range(1, 8)):   # no short self-overlap: planting must not create further copies
            continue
        found.append(s)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    recs = []
    for order in ((0, 1), (1, 0)):
        for gap in (1, 5, 19):
            left = acgt[rng.integers(0, 4, 3000 + 7 * gap)].tobytes().decode()
            mid = acgt[rng.integers(0, 4, gap)].tobytes().decode()
            right = acgt[rng.integers(0, 4, 2500)].tobytes().decode()
            recs.append((left + found[order[0]] + mid + found[order[1]] + right).encode())
    c = ctx()
    seq, off, ent = pack_batch([[r] for r in recs])
    got = c.sketch_batch(seq, off, ent, len(recs), c.params(mode="opmh", S=S, k=k, w=w))
    for e, r in enumerate(recs):
        assert np.array_equal(got["regs_u64"][e], _oracle_regs("opmh", S, k, w, [r])), e


@pytest.mark.parametrize("S,w,budget", [(512, -1, "40000"), (256, 40, "20000"), (512, -1, None)])
def test_fss_many_small_entities_long_walk_queue_in_groups(S, w, budget, monkeypatch):
    """Read-sized records with a Full SetSketch much larger than their element count: every element walks all registers and goes
    through the long-walk queue.  The queue is sized from the positions of the entities that can reach it and the entities are taken
    in groups when that exceeds the budget; registers and --save-kmers ids equal the oracle either way."""
    rng = np.random.default_rng(31 + S)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    recs = [acgt[rng.integers(0, 4, size=int(rng.choice([60, 100, 150, 151, 400])))].tobytes() for _ in range(1200)]
    if budget:
        monkeypatch.setenv("D2G_FSS_QUEUE_ELEMS", budget)
    c = ctx()
    seq, off, ent = pack_batch([[r] for r in recs])
    r = c.sketch_batch(seq, off, ent, len(recs), c.params(mode="fss", S=S, k=31, w=w), want_ids=True)
    cards, sigs = O.sketch_records_byseq(recs, "fss", S, 31, w)
    assert np.array_equal(u64(r["sig"]), u64(sigs))
    L = O.lib()
    for e in (0, 7, 500, 1199):
        hv = O.hash_stream(recs[e], 31, w)
        regs = np.empty(2 * S - 1); ids = np.zeros(S, dtype=np.uint64)
        L.d2o_css_reset(regs, S); L.d2o_css_update(regs, S, hv, len(hv), ids.ctypes.data)
        assert np.array_equal(r["ids"][e], ids), e


@pytest.mark.parametrize("n,S,shape,measure,kind", [(3000, 256, "symmetric", "similarity", 0), (2500, 100, "asymmetric", "containment", 0),
                                                     (1800, 512, "panel", "poisson_llr", 0), (2000, 128, "symmetric", "similarity", 1)])
def test_sharded_compare_single_rank_equals_plain_compare(n, S, shape, measure, kind):
    """d2g_cmp_rows_sharded_dev with a one-rank communicator walks the whole exchange path (register all-to-all, rank slices,
    all-gather, codes from the gathered ranks) and must give the float32 rows of d2g_cmp_rows / the oracle."""
    import torch
    from dashing2_b200 import capi, synth
    regs, cards = synth.synthetic_sketches(n, S, seed=n + S, n_families=max(2, n // 50), p_lo=0.02, p_hi=0.8)
    cards = cards * (1 + np.arange(n) % 4)
    if kind == 1:
        regs = np.round(regs * 64) / 64                      # equality kind: make equal registers common
    c = capi.Context(0)
    c.comm_init_rank(1, 0, c.comm_unique_id())
    nq = 700 if shape == "panel" else 0
    p = c.cmp_params(S, n, shape, measure, k=31, cmp_kind=kind, nq=nq)
    nrows = n - nq
    dev = torch.device("cuda", 0)
    t_regs = torch.from_numpy(regs).to(dev); t_cards = torch.from_numpy(cards).to(dev)
    for r0, r1 in ((0, nrows), (nrows // 3, nrows // 2)):
        nv = c.cmp_rows_size(p, r0, r1)
        out = torch.empty(nv, dtype=torch.float32, device=dev)
        c.cmp_rows_sharded_dev(p, t_regs.data_ptr(), t_cards.data_ptr(), 0, n, r0, r1, out.data_ptr())
        c.sync()
        exp = ctx().cmp_rows(regs, cards, p, r0, r1)
        assert np.array_equal(out.cpu().numpy().view(np.uint32), exp.view(np.uint32)), (shape, r0, r1)
    full = O.allpairs(regs, cards, shape, measure, k=31, cmp_kind=kind, nq=nq)
    nv = c.cmp_rows_size(p, 0, nrows)
    out = torch.empty(nv, dtype=torch.float32, device=dev)
    c.cmp_rows_sharded_dev(p, t_regs.data_ptr(), t_cards.data_ptr(), 0, n, 0, nrows, out.data_ptr())
    c.sync()
    assert np.array_equal(out.cpu().numpy().view(np.uint32), full.view(np.uint32))
    c.close()


@pytest.mark.parametrize("mode,S,k,w", [("opmh", 64, 31, -1), ("opmh", 128, 21, 30), ("fss", 64, 31, -1), ("bmh", 32, 31, -1), ("pmh", 32, 31, -1)])
def test_kmer_counts_match_reference_golden(mode, S, k, w):
    """--save-kmercounts (d2g_kmer_counts): multiplicity / weight of the element behind every register, against the float32 counts the
    unmodified reference binary wrote (tests/golden/make_golden_kmercounts.py)."""
    from dashing2_b200 import capi
    name = {("opmh", 64): "kmercounts_opmh_k31_S64", ("opmh", 128): "kmercounts_opmh_k21_w30_S128", ("fss", 64): "kmercounts_fss_k31_S64",
            ("bmh", 32): "kmercounts_bmh_k31_S32", ("pmh", 32): "kmercounts_pmh_k31_S32"}[(mode, S)]
    z = np.load(expected(name + ".npz"))
    files = ["dup.fa.gz", "g0.fa.gz", "rep.fa.gz", "adv.fa.gz"]
    recs = [O.read_fastx(os.path.join(GOLD, "inputs", f)) for f in files]
    seq, off, ent = pack_batch(recs)
    c = ctx()
    p = c.params(mode=mode, S=S, k=k, w=w)
    r = c.sketch_batch(seq, off, ent, len(recs), p, want_ids=True)
    codes, mask, nz = capi.pack_sequences([x for rr in recs for x in rr])
    counts = c.kmer_counts(codes, mask, off, ent, len(recs), p, r["ids"])
    assert np.array_equal(counts, z["counts"]), (mode, S)


@pytest.mark.parametrize("path", ["f64", "codes"])
def test_equality_compare_is_ieee_on_id_registers(path, monkeypatch):
    """The equality branch compares RegT = double with `==` even when the registers are k-mer ids viewed as doubles (cmp_core.cpp:501-506,
    count_eq.h:40-45): ids whose bits are a NaN never match -- not even themselves -- and +0 / -0 match each other.  Both compare kernels
    (and the top-k refinement) against the oracle on ids with such patterns planted in matching positions."""
    monkeypatch.setenv("D2G_CMP_PATH", path)
    rng = np.random.default_rng(9)
    n, S = 300, 64
    base = rng.integers(0, 2**63, size=(6, S), dtype=np.uint64)
    ids = base[rng.integers(0, 6, n)].copy()
    mut = rng.random((n, S)) < 0.3
    ids[mut] = rng.integers(0, 2**63, size=int(mut.sum()), dtype=np.uint64)
    ids[:, 3] = np.uint64(0x7FF8000000000001)                      # a NaN pattern everywhere in one column: never equal
    ids[::2, 5] = np.uint64(0x8000000000000000); ids[1::2, 5] = 0   # -0 and +0: equal
    ids[::3, 7] = np.uint64(0xFFF0000000000123)                     # NaN in a third of the rows
    regs = ids.view(np.float64)
    cards = rng.uniform(1e3, 1e4, n)
    c = ctx()
    for measure in ("similarity", "containment", "poisson_llr"):
        exp = O.allpairs(regs, cards, "symmetric", measure, k=31, cmp_kind=1)
        got = c.cmp_matrix(regs, cards, c.cmp_params(S, n, "symmetric", measure, k=31, cmp_kind=1))
        assert np.array_equal(got.view(np.uint32), exp.view(np.uint32)), measure
    ep = O.topk(regs, cards, 7, "similarity", k=31, cmp_kind=1)
    gp = c.lsh_topk(regs, cards, 7, "similarity", k=31, cmp_kind=1)
    assert np.array_equal(gp[0], ep[0]) and np.array_equal(gp[1], ep[1]) and np.array_equal(gp[2].view(np.uint32), ep[2].view(np.uint32))
