"""ctypes binding to oracle/libd2oracle.so (the CPU restatement; TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
from __future__ import annotations

import ctypes as C
import gzip
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = os.path.join(ROOT, "oracle", "libd2oracle.so")

u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")
f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")

MEASURES = {"similarity": 0, "containment": 1, "symmetric_containment": 2, "poisson_llr": 3,
            "intersection": 4, "union_size": 5}


def build():
    src = [os.path.join(ROOT, "oracle", f) for f in ("d2_oracle.c", "d2_oracle.h")]
    if (not os.path.exists(_LIB)) or any(os.path.getmtime(s) > os.path.getmtime(_LIB) for s in src):
        subprocess.check_call(["make", "-s", "-f", "oracle/Makefile", "oracle/libd2oracle.so"], cwd=ROOT)


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_LIB)
    L.d2o_wang64.restype = C.c_uint64; L.d2o_wang64.argtypes = [C.c_uint64]
    L.d2o_wang64_inv.restype = C.c_uint64; L.d2o_wang64_inv.argtypes = [C.c_uint64]
    L.d2o_revcomp.restype = C.c_uint64; L.d2o_revcomp.argtypes = [C.c_uint64, C.c_int]
    L.d2o_xormask_for_seed.restype = C.c_uint64; L.d2o_xormask_for_seed.argtypes = [C.c_uint64]
    L.d2o_hash_stream.restype = C.c_uint64
    L.d2o_hash_stream.argtypes = [C.c_char_p, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_uint64, u64p, C.c_uint64]
    L.d2o_hash_stream_rolling.restype = C.c_uint64
    L.d2o_hash_stream_rolling.argtypes = [C.c_char_p, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_uint64, u64p, C.c_uint64]
    L.d2o_hash_stream_protein.restype = C.c_uint64
    L.d2o_hash_stream_protein.argtypes = [C.c_char_p, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_uint64, u64p, C.c_uint64]
    L.d2o_opmh_m.restype = C.c_uint32; L.d2o_opmh_m.argtypes = [C.c_uint32]
    L.d2o_opmh_reset.argtypes = [u64p, f64p, C.c_uint32]
    L.d2o_opmh_update.argtypes = [u64p, f64p, C.c_uint32, u64p, C.c_uint64]
    L.d2o_opmh_update_mincount.argtypes = [u64p, f64p, C.c_uint32, u64p, C.c_uint64, C.c_double]
    L.d2o_opmh_card.restype = C.c_double; L.d2o_opmh_card.argtypes = [u64p, C.c_uint32]
    L.d2o_opmh_sigs.argtypes = [u64p, C.c_uint32, f64p]
    L.d2o_opmh_ids.argtypes = [u64p, C.c_uint32, u64p]
    L.d2o_css_reset.argtypes = [f64p, C.c_uint32]
    L.d2o_css_update.argtypes = [f64p, C.c_uint32, u64p, C.c_uint64, C.c_void_p]
    L.d2o_css_card.restype = C.c_double; L.d2o_css_card.argtypes = [f64p, C.c_uint32]
    L.d2o_densify.restype = C.c_uint64; L.d2o_densify.argtypes = [f64p, C.c_void_p, C.c_uint64]
    L.d2o_finalize.restype = C.c_float
    L.d2o_finalize.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int]
    for nm in ("d2o_allpairs_symmetric", "d2o_allpairs_asymmetric"):
        getattr(L, nm).argtypes = [f64p, f64p, C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_int, f32p]
    L.d2o_panel.argtypes = [f64p, f64p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_int, f32p]
    L.d2o_count_exact.restype = C.c_uint64; L.d2o_count_exact.argtypes = [u64p, C.c_uint64, u64p, f64p]
    L.d2o_count_sketch.restype = C.c_uint64; L.d2o_count_sketch.argtypes = [u64p, C.c_uint64, C.c_uint64, C.c_double, u64p, f64p]
    L.d2o_byseq_cardinality.restype = C.c_double; L.d2o_byseq_cardinality.argtypes = [C.c_double, C.c_uint64, u64p, C.c_uint64]
    L.d2o_pmh_reset.argtypes = [f64p, C.c_void_p, C.c_uint32]
    for nm in ("d2o_pmh_update", "d2o_bmh_update"):
        getattr(L, nm).restype = C.c_double
        getattr(L, nm).argtypes = [f64p, C.c_void_p, C.c_uint32, u64p, f64p, C.c_uint64, C.c_double]
    L.d2o_set_nlsh.argtypes = [C.c_int]
    L.d2o_topk_compressed.restype = C.c_uint64
    L.d2o_topk_compressed.argtypes = [f64p, f64p, f64p, C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_longdouble, u64p,
                                      C.POINTER(C.POINTER(C.c_uint32)), C.POINTER(C.POINTER(C.c_float))]
    L.d2o_nn_threshold.restype = C.c_uint64
    L.d2o_nn_threshold.argtypes = [f64p, f64p, C.c_uint64, C.c_uint64, C.c_double, C.c_int, C.c_int, C.c_int, u64p,
                                   C.POINTER(C.POINTER(C.c_uint32)), C.POINTER(C.POINTER(C.c_float))]
    L.d2o_topk.restype = C.c_uint64
    L.d2o_topk.argtypes = [f64p, f64p, C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int, u64p,
                           C.POINTER(C.POINTER(C.c_uint32)), C.POINTER(C.POINTER(C.c_float))]
    L.d2o_make_compressed.restype = C.c_int
    L.d2o_make_compressed.argtypes = [f64p, C.c_void_p, C.c_uint64, C.c_double, C.c_int, C.POINTER(C.c_longdouble), C.POINTER(C.c_longdouble), f64p]
    L.d2o_allpairs_compressed.argtypes = [f64p, f64p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int,
                                          C.c_double, C.c_longdouble, f32p]
    L.d2o_free.argtypes = [C.c_void_p]
    _lib = L
    return L


# ---------------------------------------------------------------------------------------------
# FASTA/FASTQ record reader with kseq semantics (bonsai/klib/kseq.h:178): '>' or '@' starts a
# record, the name line is dropped, sequence lines are concatenated without line terminators,
# FASTQ '+' line and qualities are skipped.
# ---------------------------------------------------------------------------------------------
def read_fastx(path: str):
    opener = gzip.open if path.endswith(".gz") else open
    with opener(path, "rb") as f:
        data = f.read()
    recs = []
    lines = data.split(b"\n")
    i, n = 0, len(lines)
    while i < n:
        ln = lines[i]
        if not ln or ln[:1] not in (b">", b"@"):
            i += 1
            continue
        fastq = ln[:1] == b"@"
        i += 1
        seq = []
        while i < n and lines[i][:1] not in (b">", b"+", b"@"):
            seq.append(lines[i].rstrip(b"\r"))
            i += 1
        s = b"".join(seq)
        if fastq and i < n and lines[i][:1] == b"+":
            i += 1
            got = 0
            while i < n and got < len(s):
                got += len(lines[i].rstrip(b"\r"))
                i += 1
        recs.append(s)
    return recs


def hash_stream(seq: bytes, k: int, w: int = -1, canon: bool = True, seed: int = 0, alphabet: int = 4) -> np.ndarray:
    L = lib()
    out = np.empty(2 * len(seq) + 4, dtype=np.uint64)
    if alphabet != 4:                                                       # protein alphabets: never canonical (src/options.h:328-331)
        n = L.d2o_hash_stream_protein(seq, len(seq), k, w, alphabet, L.d2o_xormask_for_seed(seed), out, len(out))
        return out[:n].copy()
    fn = L.d2o_hash_stream_rolling if k > 32 else L.d2o_hash_stream      # k > 32: RollingHasher (src/fastxsketch.cpp:399-410)
    n = fn(seq, len(seq), k, w, int(canon), L.d2o_xormask_for_seed(seed), out, len(out))
    return out[:n].copy()


def sketch_file(path: str, mode: str, S: int, k: int, w: int = -1, canon: bool = True, seed: int = 0,
                count_threshold: float = 0.0, cssize: int = 0, alphabet: int = 4, filterset=None):
    """Oracle equivalent of one iteration of the per-file loop (src/fastxsketch.cpp:303-624).

    Returns dict(card=..., sig=f64[S], regs_u64=..., ids=...)."""
    L = lib()
    streams = [hash_stream(r, k, w, canon, seed, alphabet) for r in read_fastx(path)]
    hv = np.concatenate(streams) if streams else np.empty(0, dtype=np.uint64)
    if filterset is not None:       # --filterset: `if(!fs_->in_set(x)) func(x)`, src/fastxsketch.cpp:385-388 (sorted hash set, src/filterset.h:207-213)
        hv = hv[~np.isin(hv, filterset)]
    if mode == "opmh":
        m = L.d2o_opmh_m(S)
        regs = np.empty(m, dtype=np.uint64); counts = np.empty(m, dtype=np.float64)
        L.d2o_opmh_reset(regs, counts, m)
        if count_threshold > 1:
            L.d2o_opmh_update_mincount(regs, counts, m, hv, len(hv), float(count_threshold))
        else:
            L.d2o_opmh_update(regs, counts, m, hv, len(hv))
        sig = np.empty(m, dtype=np.float64)
        L.d2o_opmh_sigs(regs, m, sig)
        ids = np.empty(m, dtype=np.uint64)
        L.d2o_opmh_ids(regs, m, ids)
        return dict(card=L.d2o_opmh_card(regs, m), sig=sig[:S], regs_u64=regs, ids=ids[:S],
                    counts=counts[:S], n_hashed=len(hv))
    if mode == "fss":
        regs = np.empty(2 * S - 1, dtype=np.float64)
        L.d2o_css_reset(regs, S)
        ids = np.zeros(S, dtype=np.uint64)
        L.d2o_css_update(regs, S, hv, len(hv), ids.ctypes.data)
        return dict(card=L.d2o_css_card(regs, S), sig=regs[:S].copy(), ids=ids, n_hashed=len(hv))
    if mode in ("pmh", "bmh"):
        return weighted_sketch(hv, mode, S, count_threshold, cssize)
    raise ValueError(mode)


def sketch_records_byseq(records, mode: str, S: int, k: int, w: int = -1, canon: bool = True, seed: int = 0, alphabet: int = 4):
    """Oracle equivalent of --parse-by-seq (resize_fill, src/fastxsketchbyseq.cpp:284-531): one sketch per record; set sketches
    with an estimate below 10 * S carry the exact distinct count.  Returns (cards f64[n], sigs f64[n][S])."""
    import tempfile
    L = lib()
    cards, sigs = [], []
    for rec in records:
        with tempfile.NamedTemporaryFile(suffix=".fa") as f:
            f.write(b">r\n" + rec + b"\n"); f.flush()
            o = sketch_file(f.name, mode, S, k, w, canon, seed, alphabet=alphabet)
        card = o["card"]
        if mode in ("opmh", "fss"):
            hv = hash_stream(rec, k, w, canon, seed, alphabet)
            card = L.d2o_byseq_cardinality(card, S, hv, len(hv))
        cards.append(card); sigs.append(o["sig"])
    return np.asarray(cards), np.stack(sigs) if sigs else np.empty((0, S))


def weighted_sketch(hv: np.ndarray, mode: str, S: int, count_threshold: float = 0.0, cssize: int = 0):
    """Counter (exact, or a count sketch of cssize buckets) + ProbMinHash3 / BagMinHash2 over a hashed k-mer stream
    (src/fastxsketch.cpp:429-449)."""
    L = lib()
    if cssize:
        keys = np.empty(cssize, dtype=np.uint64); cnt = np.empty(cssize, dtype=np.float64)
        nd = L.d2o_count_sketch(np.ascontiguousarray(hv), len(hv), cssize, float(count_threshold), keys, cnt)
        count_threshold = 0.0            # already applied (>=, src/counter.h:135)
    else:
        keys = np.empty(len(hv) + 1, dtype=np.uint64); cnt = np.empty(len(hv) + 1, dtype=np.float64)
        nd = L.d2o_count_exact(hv.copy(), len(hv), keys, cnt)
    regs = np.empty(2 * S - 1, dtype=np.float64)
    ids = np.zeros(S, dtype=np.uint64)
    L.d2o_pmh_reset(regs, ids.ctypes.data, S)
    fn = L.d2o_pmh_update if mode == "pmh" else L.d2o_bmh_update
    tw = fn(regs, ids.ctypes.data, S, keys[:nd].copy(), cnt[:nd].copy(), nd, float(count_threshold))
    return dict(card=tw, sig=regs[:S].copy(), ids=ids, n_hashed=len(hv), n_distinct=nd)


def densify(sig: np.ndarray) -> np.ndarray:
    out = np.ascontiguousarray(sig, dtype=np.float64).copy()
    lib().d2o_densify(out, None, len(out))
    return out


def allpairs(regs: np.ndarray, cards: np.ndarray, kind: str = "symmetric", measure: str = "similarity",
             k: int = 31, cmp_kind: int = 0, nq: int = 0) -> np.ndarray:
    L = lib()
    regs = np.ascontiguousarray(regs, dtype=np.float64); cards = np.ascontiguousarray(cards, dtype=np.float64)
    n, S = regs.shape
    me = MEASURES[measure]
    if kind == "symmetric":
        out = np.empty(n * (n - 1) // 2, dtype=np.float32)
        L.d2o_allpairs_symmetric(regs, cards, n, S, me, k, cmp_kind, out)
    elif kind == "asymmetric":
        out = np.empty(n * n, dtype=np.float32)
        L.d2o_allpairs_asymmetric(regs, cards, n, S, me, k, cmp_kind, out)
    elif kind == "panel":
        nf = n - nq
        out = np.empty(nf * nq, dtype=np.float32)
        L.d2o_panel(regs, cards, nf, nq, S, me, k, cmp_kind, out)
    else:
        raise ValueError(kind)
    return out


def make_compressed(regs: np.ndarray, fd: float, bbit: bool, a: float = -1.0, b: float = -1.0, kmers=None):
    """make_compressed (src/cmp_core.cpp:209-322): returns (quantised registers as f64, truncation used, a, b)."""
    L = lib()
    regs = np.ascontiguousarray(regs, dtype=np.float64)
    out = np.empty_like(regs)
    la = C.c_longdouble(a); lb = C.c_longdouble(b)
    kp = None if kmers is None else np.ascontiguousarray(kmers, dtype=np.uint64).ctypes.data
    trunc = L.d2o_make_compressed(regs.reshape(-1), kp, regs.size, float(fd), 1 if bbit else 0, C.byref(la), C.byref(lb), out.reshape(-1))
    return out, trunc, la, lb


def allpairs_compressed(cregs, cards, kind, measure, fd, bbit, b, k=31, nq=0):
    L = lib()
    cregs = np.ascontiguousarray(cregs, dtype=np.float64); cards = np.ascontiguousarray(cards, dtype=np.float64)
    n, S = cregs.shape
    shape = {"symmetric": 0, "asymmetric": 1, "panel": 2}[kind]
    nout = n * (n - 1) // 2 if shape == 0 else n * n if shape == 1 else (n - nq) * nq
    out = np.empty(nout, dtype=np.float32)
    L.d2o_allpairs_compressed(cregs, cards, n, nq, S, shape, MEASURES[measure], k, 1 if bbit else 0, float(fd), b, out)
    return out


def topk(regs, cards, K, measure="similarity", k=31, cmp_kind=0, nlsh=2):
    """Reference -p1 top-k pipeline -> CSR (indptr, indices, data)."""
    L = lib()
    regs = np.ascontiguousarray(regs, dtype=np.float64); cards = np.ascontiguousarray(cards, dtype=np.float64)
    n, S = regs.shape
    indptr = np.zeros(n + 1, dtype=np.uint64)
    pi = C.POINTER(C.c_uint32)(); pv = C.POINTER(C.c_float)()
    L.d2o_set_nlsh(int(nlsh))
    try:
        nnz = L.d2o_topk(regs, cards, n, S, K, MEASURES[measure], k, cmp_kind, indptr, C.byref(pi), C.byref(pv))
    finally:
        L.d2o_set_nlsh(2)
    idx = np.ctypeslib.as_array(pi, shape=(max(nnz, 1),))[:nnz].copy()
    val = np.ctypeslib.as_array(pv, shape=(max(nnz, 1),))[:nnz].copy()
    L.d2o_free(pi); L.d2o_free(pv)
    return indptr, idx, val


def topk_compressed(regs, cregs, cards, K, fd, bbit, b, measure="similarity", k=31):
    """--topk K with --fastcmp fd [--bbit-sigs]: CSR (indptr, indices, data)."""
    L = lib()
    regs = np.ascontiguousarray(regs, dtype=np.float64); cregs = np.ascontiguousarray(cregs, dtype=np.float64)
    cards = np.ascontiguousarray(cards, dtype=np.float64)
    n, S = regs.shape
    indptr = np.zeros(n + 1, dtype=np.uint64)
    pi = C.POINTER(C.c_uint32)(); pv = C.POINTER(C.c_float)()
    nnz = L.d2o_topk_compressed(regs, cregs, cards, n, S, K, MEASURES[measure], k, 1 if bbit else 0, float(fd), b, indptr, C.byref(pi), C.byref(pv))
    idx = np.ctypeslib.as_array(pi, shape=(max(nnz, 1),))[:nnz].copy()
    val = np.ctypeslib.as_array(pv, shape=(max(nnz, 1),))[:nnz].copy()
    L.d2o_free(pi); L.d2o_free(pv)
    return indptr, idx, val


def nn_threshold(regs, cards, min_sim, measure="similarity", k=31, cmp_kind=0):
    """Reference -p1 similarity-threshold graph -> CSR (indptr, indices, data)."""
    L = lib()
    regs = np.ascontiguousarray(regs, dtype=np.float64); cards = np.ascontiguousarray(cards, dtype=np.float64)
    n, S = regs.shape
    indptr = np.zeros(n + 1, dtype=np.uint64)
    pi = C.POINTER(C.c_uint32)(); pv = C.POINTER(C.c_float)()
    nnz = L.d2o_nn_threshold(regs, cards, n, S, float(min_sim), MEASURES[measure], k, cmp_kind, indptr, C.byref(pi), C.byref(pv))
    idx = np.ctypeslib.as_array(pi, shape=(max(nnz, 1),))[:nnz].copy()
    val = np.ctypeslib.as_array(pv, shape=(max(nnz, 1),))[:nnz].copy()
    L.d2o_free(pi); L.d2o_free(pv)
    return indptr, idx, val


def read_csr(path):
    raw = open(path, "rb").read()
    n, nnz = (int(x) for x in np.frombuffer(raw, np.uint64, 2))
    ip = np.frombuffer(raw, np.uint64, n + 1, 16)
    ix = np.frombuffer(raw, np.uint32, nnz, 16 + 8 * (n + 1))
    dv = np.frombuffer(raw, np.float32, nnz, 16 + 8 * (n + 1) + 4 * nnz)
    return ip, ix, dv


def filterset_from_fastx(path: str, k: int, w: int = -1, canon: bool = True, seed: int = 0) -> np.ndarray:
    """Dashing2Options::filterset for a FASTX path (src/d2.cpp:78-96): every maskfn'd k-mer / minimizer of the file, hashed with the options
    of the run; FilterSet::finalize sorts (src/filterset.cpp).  Duplicates stay (data_.size() counts them)."""
    hv = [hash_stream(r, k, w, canon, seed) for r in read_fastx(path)]
    return np.sort(np.concatenate(hv)) if hv else np.empty(0, dtype=np.uint64)


def contain(db_ids: np.ndarray, query_records, k: int, w: int = -1, canon: bool = True, seed: int = 0):
    """`dashing2 contain` for one query file (src/contain_main.cpp:33-58,218-243): every hashed k-mer of the query that is one of the
    database's sampled k-mers is counted; per reference, coverage = (sampled k-mers seen) / S and mean depth = (sum of their multiplicities)
    / (sampled k-mers seen) in uint32 integer division.  A k-mer a reference holds in two registers counts twice (kmer2ids keeps both)."""
    db_ids = np.ascontiguousarray(db_ids, dtype=np.uint64)
    nref, S = db_ids.shape
    hv = [hash_stream(r, k, w, canon, seed) for r in query_records]
    hv = np.concatenate(hv) if hv else np.empty(0, dtype=np.uint64)
    u, c = np.unique(hv, return_counts=True)
    pos = np.searchsorted(u, db_ids.reshape(-1))
    pos_c = np.minimum(pos, max(len(u) - 1, 0))
    hit = (u[pos_c] == db_ids.reshape(-1)) if len(u) else np.zeros(db_ids.size, dtype=bool)
    mult = np.where(hit, c[pos_c] if len(u) else 0, 0).reshape(nref, S).astype(np.uint64)
    matches = (mult > 0).sum(1).astype(np.uint32)
    sums = (mult.sum(1) & 0xFFFFFFFF).astype(np.uint32)
    cov = np.where(matches > 0, (1. / S) * matches, 0.).astype(np.float32)
    depth = np.where(matches > 0, sums // np.maximum(matches, 1), 0).astype(np.float32)
    return cov, depth
