"""GPU parity for the element streams (dashing2_b200/csrc/stream_kernels.cuh): k > 32 rolling hash, -C with a window, protein
alphabets -- against the reference-binary goldens and against the oracle on seeded inputs with Ns, short records, T runs."""
import os

import numpy as np
import pytest

import oracle_lib as O
from conftest import expected, GOLD
from gpu_util import ctx, pack_batch, pack_files

pytestmark = pytest.mark.gpu


def u64(a):
    return np.ascontiguousarray(a).view(np.uint64)


ROLLING = {
    "roll_opmh_k40_S128": dict(mode="opmh", S=128, k=40),
    "roll_opmh_k64_S64_nocanon": dict(mode="opmh", S=64, k=64, canon=False),
    "roll_opmh_k40_w60_S64": dict(mode="opmh", S=64, k=40, w=60),
    "roll_opmh_k33_w50_S64_nocanon": dict(mode="opmh", S=64, k=33, w=50, canon=False),
    "roll_fss_k45_S64_seed3": dict(mode="fss", S=64, k=45, seed=3),
}
ROLLING_FILES = ["g0.fa.gz", "g1.fa.gz", "dup.fa.gz", "adv.fa.gz", "reads.fq.gz"]


@pytest.mark.parametrize("case", sorted(ROLLING))
def test_rolling_hash_matches_reference_golden(case):
    """k > 32 (RollingHasher over CyclicHash, bonsai encoder.h:644-865): sketches of the reference binary, bit for bit."""
    z = np.load(expected(case + ".npz"))
    paths = [os.path.join(GOLD, "inputs", f) for f in ROLLING_FILES]
    c = ctx()
    seq, off, ent = pack_files(paths)
    r = c.sketch_batch(seq, off, ent, len(paths), c.params(**ROLLING[case]))
    for i, f in enumerate(ROLLING_FILES):
        assert np.array_equal(u64(r["sig"][i]), u64(z["sigs"][i])), (case, f)
    if ROLLING[case]["mode"] == "opmh":
        assert np.array_equal(u64(r["card"]), u64(z["cards"]))
    else:
        np.testing.assert_allclose(r["card"], z["cards"], rtol=1e-12)


def _adversarial_records(rng, k):
    """Records that exercise the N jump (i += k + 1), the 2k end rule, windows that never fill, runs of T, empty and short records."""
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)

    def rnd(n):
        return acgt[rng.integers(0, 4, n)].tobytes()
    recs = [rnd(5000), b"", rnd(k - 1), rnd(k), rnd(k + 1), rnd(2 * k), rnd(3 * k + 5)]
    b = bytearray(rnd(4000))
    for p in rng.integers(0, 4000, 25):
        b[p] = ord("N")
    recs.append(bytes(b))
    b = bytearray(rnd(3000))
    b[100:100 + 3 * k] = b"N" * (3 * k)                    # a long run of Ns
    b[2990] = ord("n"); b[1500] = ord("R")
    recs.append(bytes(b))
    recs.append(rnd(200) + b"T" * 150 + rnd(100) + b"T" * 31 + b"G" + b"T" * 32 + rnd(50) + b"t" * 64 + b"N" + b"T" * 40)
    recs.append(b"T" * 100)
    recs.append(b"N" * 50 + rnd(k + 3) + b"N" + rnd(k - 2) + b"N" + rnd(2 * k))
    recs.append(rnd(k + 2) + b"N")
    recs.append(b"N" + rnd(k + 2))
    return recs


def _oracle_sketch(recs_per_entity, mode, S, **kw):
    out = []
    for recs in recs_per_entity:
        hv = [O.hash_stream(r, kw["k"], kw.get("w", -1), kw.get("canon", True), kw.get("seed", 0), kw.get("alphabet", 4)) for r in recs]
        hv = np.concatenate(hv) if hv else np.empty(0, dtype=np.uint64)
        L = O.lib()
        if mode == "opmh":
            m = L.d2o_opmh_m(S)
            regs = np.empty(m, dtype=np.uint64); cnt = np.empty(m)
            L.d2o_opmh_reset(regs, cnt, m); L.d2o_opmh_update(regs, cnt, m, hv, len(hv))
            out.append(regs)
        elif mode == "fss":
            regs = np.empty(2 * S - 1); L.d2o_css_reset(regs, S); L.d2o_css_update(regs, S, hv, len(hv), None)
            out.append(regs[:S].copy())
        else:
            out.append(O.weighted_sketch(hv, mode, S)["sig"])
    return out


STREAM_CASES = [
    dict(k=33), dict(k=33, canon=False), dict(k=40, w=45), dict(k=40, w=45, canon=False), dict(k=64, w=90), dict(k=63), dict(k=65, seed=9),
    dict(k=128, w=130), dict(k=100, canon=False),
    dict(k=21, w=30, canon=False), dict(k=31, w=40, canon=False), dict(k=32, w=33, canon=False), dict(k=32, w=64, canon=False, seed=5),
    dict(k=5, w=100, canon=False), dict(k=31, w=300, canon=False),
]


@pytest.mark.parametrize("kw", STREAM_CASES, ids=lambda d: "_".join(f"{a}{b}" for a, b in d.items()))
def test_stream_opmh_matches_oracle_seeded(kw):
    rng = np.random.default_rng(7 + kw["k"])
    ents = [_adversarial_records(rng, kw["k"]), [b""], _adversarial_records(rng, kw["k"])[:4], _adversarial_records(rng, kw["k"])[7:]]
    c = ctx()
    seq, off, ent = pack_batch(ents)
    S = 128
    r = c.sketch_batch(seq, off, ent, len(ents), c.params(mode="opmh", S=S, **kw))
    exp = _oracle_sketch(ents, "opmh", S, **kw)
    for e in range(len(ents)):
        assert np.array_equal(r["regs_u64"][e], exp[e]), (kw, e)
    # the packed entry point and the device entry point see the same stream
    import torch
    dev = torch.device("cuda", 0)
    seq_t = torch.from_numpy(seq.copy()).to(dev); off_t = torch.from_numpy(off.astype(np.int64)).to(dev); ent_t = torch.from_numpy(ent.astype(np.int32)).to(dev)
    m = S
    regs_t = torch.empty((len(ents), m), dtype=torch.int64, device=dev)
    c.sketch_batch_dev(c.params(mode="opmh", S=S, **kw), seq_t.data_ptr(), off_t.data_ptr(), ent_t.data_ptr(), len(ent), len(ents), len(seq), regs_u64_d=regs_t.data_ptr())
    c.sync()
    assert np.array_equal(regs_t.cpu().numpy().view(np.uint64), np.stack(exp))


@pytest.mark.parametrize("mode", ["fss", "bmh", "pmh"])
@pytest.mark.parametrize("kw", [dict(k=40), dict(k=40, w=50), dict(k=21, w=30, canon=False)], ids=lambda d: "_".join(f"{a}{b}" for a, b in d.items()))
def test_stream_other_sketches_match_oracle_seeded(mode, kw):
    rng = np.random.default_rng(11)
    ents = [_adversarial_records(rng, kw["k"]), _adversarial_records(rng, kw["k"])[:3], [b"ACGT"]]
    c = ctx()
    seq, off, ent = pack_batch(ents)
    S = 64
    r = c.sketch_batch(seq, off, ent, len(ents), c.params(mode=mode, S=S, **kw))
    exp = _oracle_sketch(ents, mode, S, **kw)
    for e in range(len(ents)):
        assert np.array_equal(u64(r["sig"][e]), u64(exp[e])), (mode, kw, e)


def test_stream_opmh_count_threshold_and_distinct():
    """-m 2 and the exact distinct count over a rolling-hash stream (both go emit -> sort)."""
    rng = np.random.default_rng(5)
    base = _adversarial_records(rng, 40)
    ents = [base + base[:3], base[3:9]]
    kw = dict(k=40, w=44)
    c = ctx()
    seq, off, ent = pack_batch(ents)
    p = c.params(mode="opmh", S=64, count_threshold=2, **kw)
    r = c.sketch_batch(seq, off, ent, len(ents), p)
    L = O.lib()
    for e, recs in enumerate(ents):
        hv = np.concatenate([O.hash_stream(x, 40, 44) for x in recs])
        regs = np.empty(64, dtype=np.uint64); cnt = np.empty(64)
        L.d2o_opmh_reset(regs, cnt, 64); L.d2o_opmh_update_mincount(regs, cnt, 64, hv, len(hv), 2.0)
        assert np.array_equal(r["regs_u64"][e], regs)
        d = c.distinct_kmers(seq, off, ent, len(ents), c.params(mode="opmh", S=64, **kw))
        assert d[e] == len(np.unique(hv))


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
def test_protein_alphabets_match_reference_golden(case):
    """--protein / --protein14 / --protein6 / --protein8 with --parse-by-seq (the only mode in which the reference binary sketches protein
    input at all): registers bit for bit, cardinalities with the exact-count rule below 10 * S."""
    kw = PROTEIN[case]; S = kw["S"]
    z = np.load(expected(case + ".npz"))
    recs = O.read_fastx(os.path.join(GOLD, "inputs", "prot.fa.gz"))
    c = ctx()
    seq, off, ent = pack_batch([[r] for r in recs])
    p = c.params(canon=False, **kw)
    r = c.sketch_batch(seq, off, ent, len(recs), p)
    assert np.array_equal(u64(r["sig"]), u64(z["byseq_sigs"]))
    card = r["card"].copy()
    card[np.isnan(card)] = 0.
    small = np.flatnonzero(card < 10 * S)
    if len(small):
        sseq, soff, sent = pack_batch([[recs[i]] for i in small])
        card[small] = c.distinct_kmers(sseq, soff, sent, len(small), p)
    if kw["mode"] == "opmh":
        assert np.array_equal(card, z["byseq_cards"])
    else:
        np.testing.assert_allclose(card, z["byseq_cards"], rtol=1e-12)


def test_stream_limits_fail_loudly():
    from dashing2_b200.capi import D2GError
    c = ctx()
    seq, off, ent = pack_batch([[b"ACGT" * 100]])
    with pytest.raises(D2GError):
        c.sketch_batch(seq, off, ent, 1, c.params(mode="opmh", S=64, k=15, alphabet=20, canon=False))      # beyond the exact protein encoding
    with pytest.raises(D2GError):
        c.sketch_batch(seq, off, ent, 1, c.params(mode="opmh", S=64, k=7, alphabet=20, canon=True))        # protein is never canonical
    with pytest.raises(D2GError):
        c.sketch_batch(seq, off, ent, 1, c.params(mode="opmh", S=64, k=40, w=40 + 2000))                     # window beyond the tile halo


FILTERSET = {
    "fs_opmh_k31_S128": dict(mode="opmh", S=128, k=31),
    "fs_opmh_k21_w30_S64": dict(mode="opmh", S=64, k=21, w=30),
    "fs_fss_k31_S64": dict(mode="fss", S=64, k=31),
    "fs_opmh_k40_S64": dict(mode="opmh", S=64, k=40),
}


@pytest.mark.parametrize("case", sorted(FILTERSET))
def test_filterset_matches_reference_golden(case):
    """--filterset dup.fa: the set is built on the device from the filter file's records, the sketch kernels (exact and element-stream
    flavours; the 32-bit-key fast kernel steps aside) test membership in front of the consumer.  Registers of the reference binary."""
    kw = FILTERSET[case]
    z = np.load(expected(case + ".npz"))
    c = ctx()
    p = c.params(**kw)
    fseq, foff, _ = pack_files([os.path.join(GOLD, "inputs", "dup.fa.gz")])
    try:
        nfs = c.set_filterset(fseq, foff, p)
        assert nfs == len(O.filterset_from_fastx(os.path.join(GOLD, "inputs", "dup.fa.gz"), kw["k"], kw.get("w", -1)))
        paths = [os.path.join(GOLD, "inputs", f) for f in ["g0.fa.gz", "g1.fa.gz", "dup.fa.gz", "adv.fa.gz"]]
        seq, off, ent = pack_files(paths)
        r = c.sketch_batch(seq, off, ent, len(paths), p)
        assert np.array_equal(u64(r["sig"]), u64(z["sigs"]))
        if kw["mode"] == "opmh":
            assert np.array_equal(r["card"], z["cards"])
        else:
            np.testing.assert_allclose(r["card"], z["cards"], rtol=1e-12)
        # raw hashed values instead of a FASTX source: the same set, the same registers
        c.set_filterset_values(O.filterset_from_fastx(os.path.join(GOLD, "inputs", "dup.fa.gz"), kw["k"], kw.get("w", -1))[::-1].copy())
        r2 = c.sketch_batch(seq, off, ent, len(paths), p)
        assert np.array_equal(u64(r2["sig"]), u64(z["sigs"]))
    finally:
        c.clear_filterset()
    # and without the filter the registers differ (the filter did something)
    r3 = c.sketch_batch(seq, off, ent, len(paths), p)
    assert not np.array_equal(u64(r3["sig"]), u64(z["sigs"]))


@pytest.mark.parametrize("kw", [dict(k=40), dict(k=40, w=48), dict(k=21, w=30, canon=False)], ids=lambda d: "_".join(f"{a}{b}" for a, b in d.items()))
def test_stream_save_kmers_ids_and_counts(kw):
    """--save-kmers ids and -N counts over an element stream: the ids are hashed values of the stream and each count is the multiplicity
    of its id in it (one occurrence per window when w > k)."""
    from dashing2_b200 import capi
    rng = np.random.default_rng(21)
    base = _adversarial_records(rng, kw["k"])
    ents = [base + base[:2], base[2:8] + base[2:4]]
    seq, off, ent = pack_batch(ents)
    c = ctx()
    p = c.params(mode="opmh", S=64, **kw)
    r = c.sketch_batch(seq, off, ent, len(ents), p, want_ids=True)
    codes, mask, nz = capi.pack_sequences([x for rr in ents for x in rr])
    counts = c.kmer_counts(codes, mask, off, ent, len(ents), p, r["ids"])
    for e, recs in enumerate(ents):
        hv = np.concatenate([O.hash_stream(x, kw["k"], kw.get("w", -1), kw.get("canon", True)) for x in recs])
        u, cnt = np.unique(hv, return_counts=True)
        mult = dict(zip(u.tolist(), cnt.tolist()))
        filled = r["regs_u64"][e][:64] != np.uint64(0xFFFFFFFFFFFFFFFF)
        assert all(int(x) in mult for x in r["ids"][e][filled])
        exp = np.array([mult.get(int(x), 0) for x in r["ids"][e]], dtype=np.float32)
        assert np.array_equal(counts[e][filled], exp[filled]), (kw, e)
