"""ctypes binding of libd2gpu.so (include/d2gpu.h) -- the only way Python reaches the product.

The library is built in-tree (``make`` / ``__graft_entry__.build()``) as dashing2_b200/libd2gpu.so.
There is no CPU fallback anywhere in this module: if the shared library is missing ``load()`` raises,
and without a CUDA device ``Context()`` raises (D2G_ENODEVICE).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libd2gpu.so")

MODE = {"opmh": 0, "fss": 1, "bmh": 2, "pmh": 3}
MEASURE = {"similarity": 0, "containment": 1, "symmetric_containment": 2, "poisson_llr": 3,
           "intersection": 4, "union_size": 5}
SHAPE = {"symmetric": 0, "asymmetric": 1, "panel": 2}


class SketchParams(C.Structure):
    _fields_ = [("k", C.c_int32), ("w", C.c_int32), ("canon", C.c_int32), ("mode", C.c_int32),
                ("xormask", C.c_uint64), ("sketchsize", C.c_uint32), ("count_threshold", C.c_uint32),
                ("countsketch_size", C.c_uint64), ("alphabet", C.c_int32), ("reserved", C.c_int32)]


class CmpParams(C.Structure):
    _fields_ = [("sketchsize", C.c_uint32), ("cmp_kind", C.c_int32), ("measure", C.c_int32), ("k", C.c_int32),
                ("shape", C.c_int32), ("n", C.c_uint64), ("nq", C.c_uint64), ("regbytes", C.c_double),
                ("compressed_b", C.c_longdouble), ("nlsh", C.c_int32)]


SINK_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_float), C.c_uint64, C.c_uint64, C.c_uint64)

# every symbol include/d2gpu.h declares
EXPORTS = ["d2g_init", "d2g_destroy", "d2g_last_error", "d2g_version", "d2g_stream", "d2g_sync", "d2g_launch_count",
           "d2g_set_timing", "d2g_get_timing", "d2g_stat",
           "d2g_opmh_m", "d2g_count_kmers", "d2g_sketch_batch", "d2g_distinct_kmers", "d2g_opmh_finalize", "d2g_sketch_batch_dev",
           "d2g_init_devices", "d2g_comm_unique_id", "d2g_comm_init_rank", "d2g_comm_init_all", "d2g_comm_size", "d2g_comm_rank", "d2g_comm_destroy",
           "d2g_cmp_rows_sharded_dev", "d2g_cmp_stream_sharded", "d2g_cmp_rows_sharded",
           "d2g_set_filterset", "d2g_set_filterset_values", "d2g_clear_filterset", "d2g_kmer_counts", "d2g_packed_words", "d2g_pack_sequences", "d2g_pack_dev", "d2g_sketch_batch_packed", "d2g_sketch_batch_packed_dev",
           "d2g_densify", "d2g_densify_dev", "d2g_make_compressed", "d2g_cmp_output_size", "d2g_cmp_rows_size", "d2g_cmp_matrix",
           "d2g_cmp_stream", "d2g_cmp_rows", "d2g_cmp_rows_dev", "d2g_cmp_counts", "d2g_lsh_topk", "d2g_lsh_topk_rows", "d2g_lsh_graph", "d2g_free"]

_lib = None


u64_t = C.c_uint64


class D2GError(RuntimeError):
    pass


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise D2GError(f"{LIB_PATH} not built: run `make` (or __graft_entry__.build()); there is no CPU fallback")
    L = C.CDLL(LIB_PATH)
    vp, u64, u32, i32 = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int32
    L.d2g_init.argtypes = [C.POINTER(vp), C.c_int]; L.d2g_init.restype = C.c_int
    L.d2g_destroy.argtypes = [vp]; L.d2g_destroy.restype = None
    L.d2g_last_error.restype = C.c_char_p
    L.d2g_version.restype = C.c_char_p
    L.d2g_stream.argtypes = [vp]; L.d2g_stream.restype = vp
    L.d2g_sync.argtypes = [vp]; L.d2g_sync.restype = C.c_int
    L.d2g_launch_count.argtypes = [vp]; L.d2g_launch_count.restype = u64
    L.d2g_set_timing.argtypes = [vp, C.c_int]; L.d2g_set_timing.restype = C.c_int
    L.d2g_get_timing.argtypes = [vp, C.c_int, C.POINTER(C.c_double), C.POINTER(u64)]; L.d2g_get_timing.restype = C.c_int
    L.d2g_opmh_m.argtypes = [u32]; L.d2g_opmh_m.restype = u32
    L.d2g_count_kmers.argtypes = [vp, u64, i32]; L.d2g_count_kmers.restype = u64
    L.d2g_sketch_batch.argtypes = [vp, C.POINTER(SketchParams), vp, vp, vp, u64, u32, vp, vp, vp, vp, C.POINTER(u64)]
    L.d2g_sketch_batch.restype = C.c_int
    L.d2g_opmh_finalize.argtypes = [vp, u32, u32, vp, vp]; L.d2g_opmh_finalize.restype = C.c_int
    L.d2g_distinct_kmers.argtypes = [vp, C.POINTER(SketchParams), vp, vp, vp, u64, u32, vp]; L.d2g_distinct_kmers.restype = C.c_int
    L.d2g_sketch_batch_dev.argtypes = [vp, C.POINTER(SketchParams), vp, vp, vp, u64, u32, u64, vp, vp, vp, vp]
    L.d2g_sketch_batch_dev.restype = C.c_int
    L.d2g_init_devices.argtypes = [C.POINTER(vp), C.POINTER(C.c_int), C.c_int]; L.d2g_init_devices.restype = C.c_int
    L.d2g_comm_unique_id.argtypes = [vp]; L.d2g_comm_unique_id.restype = C.c_int
    L.d2g_comm_init_rank.argtypes = [vp, C.c_int, C.c_int, vp]; L.d2g_comm_init_rank.restype = C.c_int
    L.d2g_comm_init_all.argtypes = [C.POINTER(vp), C.c_int]; L.d2g_comm_init_all.restype = C.c_int
    L.d2g_comm_size.argtypes = [vp]; L.d2g_comm_size.restype = C.c_int
    L.d2g_comm_rank.argtypes = [vp]; L.d2g_comm_rank.restype = C.c_int
    L.d2g_comm_destroy.argtypes = [vp]; L.d2g_comm_destroy.restype = C.c_int
    L.d2g_cmp_rows_sharded_dev.argtypes = [vp, C.POINTER(CmpParams), vp, vp, u64, u64, u64, u64, vp]; L.d2g_cmp_rows_sharded_dev.restype = C.c_int
    L.d2g_cmp_stream_sharded.argtypes = [vp, C.POINTER(CmpParams), vp, vp, u64, u64, u64, u64, SINK_FN, vp]; L.d2g_cmp_stream_sharded.restype = C.c_int
    L.d2g_kmer_counts.argtypes = [vp, C.POINTER(SketchParams), vp, vp, vp, vp, u64, u32, vp, vp]; L.d2g_kmer_counts.restype = C.c_int
    L.d2g_packed_words.argtypes = [u64]; L.d2g_packed_words.restype = u64
    L.d2g_pack_sequences.argtypes = [vp, vp, u64, vp, vp, C.POINTER(u64)]; L.d2g_pack_sequences.restype = C.c_int
    L.d2g_pack_dev.argtypes = [vp, vp, u64, vp, vp]; L.d2g_pack_dev.restype = C.c_int
    L.d2g_sketch_batch_packed.argtypes = [vp, C.POINTER(SketchParams), vp, vp, vp, vp, u64, u32, vp, vp, vp, vp, C.POINTER(u64)]
    L.d2g_sketch_batch_packed.restype = C.c_int
    L.d2g_sketch_batch_packed_dev.argtypes = [vp, C.POINTER(SketchParams), vp, vp, vp, vp, u64, u32, u64, vp, vp, vp, vp]
    L.d2g_sketch_batch_packed_dev.restype = C.c_int
    L.d2g_densify.argtypes = [vp, vp, vp, u64, u32]; L.d2g_densify.restype = C.c_int
    L.d2g_densify_dev.argtypes = [vp, vp, vp, u64, u32]; L.d2g_densify_dev.restype = C.c_int
    L.d2g_make_compressed.argtypes = [vp, vp, u64, u32, C.c_double, i32, C.POINTER(C.c_longdouble), C.POINTER(C.c_longdouble), vp, C.POINTER(i32)]
    L.d2g_make_compressed.restype = C.c_int
    L.d2g_cmp_output_size.argtypes = [C.POINTER(CmpParams)]; L.d2g_cmp_output_size.restype = u64
    L.d2g_cmp_rows_size.argtypes = [C.POINTER(CmpParams), u64, u64, C.POINTER(u64)]; L.d2g_cmp_rows_size.restype = C.c_int
    L.d2g_cmp_matrix.argtypes = [vp, C.POINTER(CmpParams), vp, vp, vp]; L.d2g_cmp_matrix.restype = C.c_int
    L.d2g_cmp_stream.argtypes = [vp, C.POINTER(CmpParams), vp, vp, u64, u64, SINK_FN, vp]; L.d2g_cmp_stream.restype = C.c_int
    L.d2g_cmp_rows.argtypes = [vp, C.POINTER(CmpParams), vp, vp, u64, u64, vp]; L.d2g_cmp_rows.restype = C.c_int
    L.d2g_cmp_rows_dev.argtypes = [vp, C.POINTER(CmpParams), vp, vp, u64, u64, vp]; L.d2g_cmp_rows_dev.restype = C.c_int
    L.d2g_cmp_counts.argtypes = [vp, u32, i32, vp, u64, vp, u64, vp, vp]; L.d2g_cmp_counts.restype = C.c_int
    L.d2g_lsh_topk.argtypes = [vp, C.POINTER(CmpParams), vp, vp, i32, vp, C.POINTER(C.POINTER(C.c_uint32)), C.POINTER(C.POINTER(C.c_float))]
    L.d2g_lsh_topk.restype = C.c_int
    L.d2g_lsh_topk_rows.argtypes = [vp, C.POINTER(CmpParams), vp, vp, i32, u64, u64, vp, C.POINTER(C.POINTER(C.c_uint32)), C.POINTER(C.POINTER(C.c_float))]
    L.d2g_lsh_topk_rows.restype = C.c_int
    L.d2g_lsh_graph.argtypes = [vp, C.POINTER(CmpParams), vp, vp, vp, i32, C.c_double, u64, u64, vp, C.POINTER(C.POINTER(C.c_uint32)), C.POINTER(C.POINTER(C.c_float))]
    L.d2g_lsh_graph.restype = C.c_int
    L.d2g_cmp_rows_sharded.argtypes = [vp, C.POINTER(CmpParams), vp, vp, u64, u64, u64, u64, vp]; L.d2g_cmp_rows_sharded.restype = C.c_int
    L.d2g_stat.argtypes = [vp, C.c_int]; L.d2g_stat.restype = u64
    L.d2g_set_filterset.argtypes = [vp, C.POINTER(SketchParams), vp, vp, u64, C.POINTER(u64)]; L.d2g_set_filterset.restype = C.c_int
    L.d2g_set_filterset_values.argtypes = [vp, vp, u64]; L.d2g_set_filterset_values.restype = C.c_int
    L.d2g_clear_filterset.argtypes = [vp]; L.d2g_clear_filterset.restype = C.c_int
    L.d2g_free.argtypes = [vp]; L.d2g_free.restype = None
    _lib = L
    return L


def _check(rc):
    if rc != 0:
        raise D2GError(f"libd2gpu error {rc}: {load().d2g_last_error().decode(errors='replace')}")


def _ptr(a):
    return None if a is None else a.ctypes.data


def xormask_for_seed(seed: int) -> int:
    """maskfn XOR mask (src/enums.cpp:133-140): 0 for seed 0, else Wang(seed)."""
    if seed == 0:
        return 0
    M = (1 << 64) - 1
    key = seed & M
    key = (~key + (key << 21)) & M
    key ^= key >> 24
    key = (key + (key << 3) + (key << 8)) & M
    key ^= key >> 14
    key = (key + (key << 2) + (key << 4)) & M
    key ^= key >> 28
    key = (key + (key << 31)) & M
    return key


def pack_sequences(pieces):
    """Host packer (d2g_pack_sequences): list of bytes-like ASCII pieces -> (codes u64[], mask u32[], n_invalid_words).
    Needs no device."""
    L = load()
    bufs = [np.frombuffer(bytes(x), dtype=np.uint8) if not isinstance(x, np.ndarray) else np.ascontiguousarray(x, dtype=np.uint8) for x in pieces]
    n = len(bufs)
    ptrs = (C.c_void_p * max(n, 1))(*[b.ctypes.data if b.size else None for b in bufs])
    lens = np.asarray([b.size for b in bufs], dtype=np.uint64)
    nw = int(L.d2g_packed_words(int(lens.sum())))
    codes = np.empty(nw, dtype=np.uint64); mask = np.empty(nw, dtype=np.uint32)
    nz = C.c_uint64(0)
    _check(L.d2g_pack_sequences(ptrs, _ptr(lens), n, _ptr(codes), _ptr(mask), C.byref(nz)))
    return codes, mask, int(nz.value)


class Context:
    """One libd2gpu context (device + stream + scratch)."""

    def __init__(self, device: int = 0):
        self.L = load()
        h = C.c_void_p()
        _check(self.L.d2g_init(C.byref(h), device))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.L.d2g_destroy(self.h)
            self.h = None

    __del__ = close

    @property
    def stream(self) -> int:
        return self.L.d2g_stream(self.h) or 0

    def sync(self):
        _check(self.L.d2g_sync(self.h))

    def launch_count(self) -> int:
        return int(self.L.d2g_launch_count(self.h))

    def stat(self, which: int = 0) -> int:
        return int(self.L.d2g_stat(self.h, which))

    def set_timing(self, on: bool):
        _check(self.L.d2g_set_timing(self.h, int(on)))

    def get_timing(self, kernel_class: int):
        """(total ms, launches) of one kernel class since the last call; 0 sketch main, 1 sketch boot, 2 cmp."""
        ms = C.c_double(0); n = C.c_uint64(0)
        _check(self.L.d2g_get_timing(self.h, kernel_class, C.byref(ms), C.byref(n)))
        return float(ms.value), int(n.value)

    # ---- sketch ----
    @staticmethod
    def params(mode="opmh", S=1024, k=31, w=-1, canon=True, seed=0, count_threshold=0, cssize=0, alphabet=0):
        return SketchParams(k, w, int(canon), MODE[mode], xormask_for_seed(seed), S, int(count_threshold), int(cssize), int(alphabet), 0)

    def sketch_batch(self, seq: np.ndarray, rec_off: np.ndarray, rec_entity: np.ndarray, n_entities: int,
                     p: SketchParams, want_ids=False, want_regs=True):
        """Host buffers in, host arrays out (d2g_sketch_batch)."""
        seq = np.ascontiguousarray(seq, dtype=np.uint8)
        rec_off = np.ascontiguousarray(rec_off, dtype=np.uint64)
        rec_entity = np.ascontiguousarray(rec_entity, dtype=np.uint32)
        n_rec = len(rec_entity)
        S = p.sketchsize
        m = self.L.d2g_opmh_m(S)
        regs = np.empty((n_entities, m), dtype=np.uint64) if (p.mode == 0 and want_regs) else None
        sig = np.empty((n_entities, S), dtype=np.float64)
        card = np.empty(n_entities, dtype=np.float64)
        ids = np.empty((n_entities, S), dtype=np.uint64) if want_ids else None
        nk = C.c_uint64(0)
        _check(self.L.d2g_sketch_batch(self.h, C.byref(p), _ptr(seq), _ptr(rec_off), _ptr(rec_entity), n_rec, n_entities,
                                       _ptr(regs), _ptr(sig), _ptr(card), _ptr(ids), C.byref(nk)))
        return dict(regs_u64=regs, sig=sig, card=card, ids=ids, n_kmers=int(nk.value))

    def sketch_batch_packed(self, codes: np.ndarray, mask, rec_off: np.ndarray, rec_entity: np.ndarray, n_entities: int,
                            p: SketchParams, want_ids=False, want_regs=True):
        """The same with the sequence packed by the caller (d2g_sketch_batch_packed); mask may be None."""
        codes = np.ascontiguousarray(codes, dtype=np.uint64)
        mask = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint32)
        rec_off = np.ascontiguousarray(rec_off, dtype=np.uint64)
        rec_entity = np.ascontiguousarray(rec_entity, dtype=np.uint32)
        S = p.sketchsize
        m = self.L.d2g_opmh_m(S)
        regs = np.empty((n_entities, m), dtype=np.uint64) if (p.mode == 0 and want_regs) else None
        sig = np.empty((n_entities, S), dtype=np.float64)
        card = np.empty(n_entities, dtype=np.float64)
        ids = np.empty((n_entities, S), dtype=np.uint64) if want_ids else None
        nk = C.c_uint64(0)
        _check(self.L.d2g_sketch_batch_packed(self.h, C.byref(p), _ptr(codes), _ptr(mask), _ptr(rec_off), _ptr(rec_entity), len(rec_entity),
                                              n_entities, _ptr(regs), _ptr(sig), _ptr(card), _ptr(ids), C.byref(nk)))
        return dict(regs_u64=regs, sig=sig, card=card, ids=ids, n_kmers=int(nk.value))

    def kmer_counts(self, codes, mask, rec_off, rec_entity, n_entities, p: SketchParams, ids: np.ndarray) -> np.ndarray:
        """--save-kmercounts: multiplicity of the element behind every register (d2g_kmer_counts), float32 [n_entities][S]."""
        codes = np.ascontiguousarray(codes, dtype=np.uint64)
        mask = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint32)
        rec_off = np.ascontiguousarray(rec_off, dtype=np.uint64); rec_entity = np.ascontiguousarray(rec_entity, dtype=np.uint32)
        ids = np.ascontiguousarray(ids, dtype=np.uint64)
        out = np.empty((n_entities, p.sketchsize), dtype=np.float32)
        _check(self.L.d2g_kmer_counts(self.h, C.byref(p), _ptr(codes), _ptr(mask), _ptr(rec_off), _ptr(rec_entity), len(rec_entity), n_entities, _ptr(ids), _ptr(out)))
        return out

    def distinct_kmers(self, seq: np.ndarray, rec_off: np.ndarray, rec_entity: np.ndarray, n_entities: int, p: SketchParams) -> np.ndarray:
        """Exact distinct k-mers (minimizers) per entity, host in / host out (d2g_distinct_kmers)."""
        seq = np.ascontiguousarray(seq, dtype=np.uint8)
        rec_off = np.ascontiguousarray(rec_off, dtype=np.uint64)
        rec_entity = np.ascontiguousarray(rec_entity, dtype=np.uint32)
        out = np.empty(n_entities, dtype=np.uint64)
        _check(self.L.d2g_distinct_kmers(self.h, C.byref(p), _ptr(seq), _ptr(rec_off), _ptr(rec_entity), len(rec_entity), n_entities, _ptr(out)))
        return out

    def sketch_batch_dev(self, p, seq_d, rec_off_d, rec_entity_d, n_rec, n_entities, total_len,
                         regs_u64_d=0, sig_d=0, card_d=0, ids_d=0):
        """Device pointers (ints) in/out; asynchronous on self.stream (d2g_sketch_batch_dev)."""
        _check(self.L.d2g_sketch_batch_dev(self.h, C.byref(p), seq_d, rec_off_d, rec_entity_d, n_rec, n_entities, total_len,
                                           regs_u64_d or None, sig_d or None, card_d or None, ids_d or None))

    def pack_dev(self, seq_d, total_len, codes_d, mask_d):
        _check(self.L.d2g_pack_dev(self.h, seq_d, total_len, codes_d, mask_d))

    def sketch_batch_packed_dev(self, p, codes_d, mask_d, rec_off_d, rec_entity_d, n_rec, n_entities, total_len,
                                regs_u64_d=0, sig_d=0, card_d=0, ids_d=0):
        _check(self.L.d2g_sketch_batch_packed_dev(self.h, C.byref(p), codes_d, mask_d, rec_off_d, rec_entity_d, n_rec, n_entities, total_len,
                                                  regs_u64_d or None, sig_d or None, card_d or None, ids_d or None))

    def opmh_finalize(self, regs_u64: np.ndarray, S: int):
        regs_u64 = np.ascontiguousarray(regs_u64, dtype=np.uint64)
        n = regs_u64.shape[0]
        sig = np.empty((n, S), dtype=np.float64); card = np.empty(n, dtype=np.float64)
        _check(self.L.d2g_opmh_finalize(_ptr(regs_u64), n, S, _ptr(sig), _ptr(card)))
        return sig, card

    # ---- compare ----
    @staticmethod
    def cmp_params(S, n, shape="symmetric", measure="similarity", k=31, cmp_kind=0, nq=0, regbytes=8.0, compressed_b=0.0):
        return CmpParams(S, cmp_kind, MEASURE[measure], k, SHAPE[shape], n, nq, regbytes, compressed_b)

    def make_compressed(self, regs: np.ndarray, regbytes: float, bbit: bool, a=-1.0, b=-1.0, kmers: np.ndarray | None = None):
        """make_compressed (--fastcmp N [--bbit-sigs]): returns (quantised registers f64, cmp_kind to compare them with, a, b);
        a and b are ctypes.c_longdouble -- hand b to cmp_params(compressed_b=b) as is (its .value is only a double)."""
        regs = np.ascontiguousarray(regs, dtype=np.float64)
        n, S = regs.shape
        out = np.empty_like(regs)
        la = C.c_longdouble(a); lb = C.c_longdouble(b); used = C.c_int32(0)
        if kmers is not None:
            kmers = np.ascontiguousarray(kmers, dtype=np.uint64)
        _check(self.L.d2g_make_compressed(_ptr(regs), _ptr(kmers), n, S, float(regbytes), int(bool(bbit)), C.byref(la), C.byref(lb), _ptr(out), C.byref(used)))
        return out, (3 if used.value else 2), la, lb

    def densify(self, sig: np.ndarray, kmers: np.ndarray | None = None):
        sig = np.ascontiguousarray(sig, dtype=np.float64).copy()
        n, S = sig.shape
        _check(self.L.d2g_densify(self.h, _ptr(sig), _ptr(kmers), n, S))
        return sig

    def cmp_matrix(self, regs: np.ndarray, cards: np.ndarray, p: CmpParams) -> np.ndarray:
        regs = np.ascontiguousarray(regs, dtype=np.float64); cards = np.ascontiguousarray(cards, dtype=np.float64)
        out = np.empty(int(self.L.d2g_cmp_output_size(C.byref(p))), dtype=np.float32)
        _check(self.L.d2g_cmp_matrix(self.h, C.byref(p), _ptr(regs), _ptr(cards), _ptr(out)))
        return out

    def cmp_stream(self, regs, cards, p: CmpParams, row_begin, row_end, callback):
        regs = np.ascontiguousarray(regs, dtype=np.float64); cards = np.ascontiguousarray(cards, dtype=np.float64)

        def _sink(user, block, first_row, n_rows, n_vals):
            arr = np.ctypeslib.as_array(block, shape=(int(n_vals),))
            return int(callback(arr, int(first_row), int(n_rows)) or 0)
        fn = SINK_FN(_sink)
        _check(self.L.d2g_cmp_stream(self.h, C.byref(p), _ptr(regs), _ptr(cards), row_begin, row_end, fn, None))

    def cmp_rows(self, regs, cards, p: CmpParams, row_begin, row_end, out: np.ndarray | None = None) -> np.ndarray:
        """Rows [row_begin,row_end) host in / host out (d2g_cmp_rows); `out` may be a caller (pinned) float32 buffer."""
        regs = np.ascontiguousarray(regs, dtype=np.float64); cards = np.ascontiguousarray(cards, dtype=np.float64)
        nv = self.cmp_rows_size(p, row_begin, row_end)
        if out is None:
            out = np.empty(nv, dtype=np.float32)
        assert out.dtype == np.float32 and out.size >= nv and out.flags.c_contiguous
        _check(self.L.d2g_cmp_rows(self.h, C.byref(p), _ptr(regs), _ptr(cards), row_begin, row_end, _ptr(out)))
        return out[:nv]

    def cmp_rows_dev(self, p: CmpParams, regs_d, cards_d, row_begin, row_end, out_d):
        _check(self.L.d2g_cmp_rows_dev(self.h, C.byref(p), regs_d, cards_d, row_begin, row_end, out_d))

    # ---- several GPUs ----
    def comm_unique_id(self) -> bytes:
        buf = C.create_string_buffer(128)
        _check(self.L.d2g_comm_unique_id(buf))
        return buf.raw

    def comm_init_rank(self, nranks: int, rank: int, uid: bytes):
        _check(self.L.d2g_comm_init_rank(self.h, nranks, rank, C.create_string_buffer(uid, 128)))

    def cmp_rows_sharded_dev(self, p: CmpParams, local_regs_d, local_cards_d, local_begin, local_n, row_begin, row_end, out_d):
        """Collective over the context's communicator (d2g_cmp_rows_sharded_dev)."""
        _check(self.L.d2g_cmp_rows_sharded_dev(self.h, C.byref(p), local_regs_d, local_cards_d, local_begin, local_n, row_begin, row_end, out_d))

    def cmp_rows_size(self, p: CmpParams, row_begin, row_end) -> int:
        nv = C.c_uint64(0)
        _check(self.L.d2g_cmp_rows_size(C.byref(p), row_begin, row_end, C.byref(nv)))
        return int(nv.value)

    def cmp_counts(self, rows: np.ndarray, cols: np.ndarray, cmp_kind=0):
        rows = np.ascontiguousarray(rows, dtype=np.float64); cols = np.ascontiguousarray(cols, dtype=np.float64)
        nr, S = rows.shape; nc = cols.shape[0]
        c0 = np.empty((nr, nc), dtype=np.uint32); c1 = np.zeros((nr, nc), dtype=np.uint32)
        _check(self.L.d2g_cmp_counts(self.h, S, cmp_kind, _ptr(rows), nr, _ptr(cols), nc, _ptr(c0), _ptr(c1)))
        return c0, c1

    def lsh_topk(self, regs: np.ndarray, cards: np.ndarray, topk: int, measure="similarity", k=31, cmp_kind=0, rows=None, nlsh=0):
        """--topk neighbour graph (d2g_lsh_topk / d2g_lsh_topk_rows for rows=(begin, end)): returns CSR
        (indptr u64[rows+1], indices u32[nnz], data f32[nnz])."""
        regs = np.ascontiguousarray(regs, dtype=np.float64); cards = np.ascontiguousarray(cards, dtype=np.float64)
        n, S = regs.shape
        p = self.cmp_params(S, n, "symmetric", measure, k=k, cmp_kind=cmp_kind)
        p.nlsh = nlsh
        x0, x1 = rows if rows is not None else (0, n)
        indptr = np.zeros(x1 - x0 + 1, dtype=np.uint64)
        pi = C.POINTER(C.c_uint32)(); pv = C.POINTER(C.c_float)()
        _check(self.L.d2g_lsh_topk_rows(self.h, C.byref(p), _ptr(regs), _ptr(cards), topk, x0, x1, _ptr(indptr), C.byref(pi), C.byref(pv)))
        nnz = int(indptr[-1])
        idx = np.ctypeslib.as_array(pi, shape=(max(nnz, 1),))[:nnz].copy()
        val = np.ctypeslib.as_array(pv, shape=(max(nnz, 1),))[:nnz].copy()
        self.L.d2g_free(pi); self.L.d2g_free(pv)
        return indptr, idx, val

    def lsh_graph(self, regs: np.ndarray, cards: np.ndarray, topk: int = -1, min_similarity: float = 0., measure="similarity", k=31, cmp_kind=0,
                  rows=None, nlsh=0, index_regs=None, regbytes=8.0, compressed_b=0.0):
        """d2g_lsh_graph: top-k lists (topk > 0) or a similarity-threshold graph (topk <= 0); index_regs = the f64 signatures when regs
        are compressed registers (--topk with --fastcmp).  Returns CSR (indptr, indices, data)."""
        regs = np.ascontiguousarray(regs, dtype=np.float64); cards = np.ascontiguousarray(cards, dtype=np.float64)
        if index_regs is not None:
            index_regs = np.ascontiguousarray(index_regs, dtype=np.float64)
        n, S = regs.shape
        p = self.cmp_params(S, n, "symmetric", measure, k=k, cmp_kind=cmp_kind, regbytes=regbytes, compressed_b=compressed_b)
        p.nlsh = nlsh
        x0, x1 = rows if rows is not None else (0, n)
        indptr = np.zeros(x1 - x0 + 1, dtype=np.uint64)
        pi = C.POINTER(C.c_uint32)(); pv = C.POINTER(C.c_float)()
        _check(self.L.d2g_lsh_graph(self.h, C.byref(p), _ptr(index_regs), _ptr(regs), _ptr(cards), int(topk), float(min_similarity), x0, x1,
                                    _ptr(indptr), C.byref(pi), C.byref(pv)))
        nnz = int(indptr[-1])
        idx = np.ctypeslib.as_array(pi, shape=(max(nnz, 1),))[:nnz].copy()
        val = np.ctypeslib.as_array(pv, shape=(max(nnz, 1),))[:nnz].copy()
        self.L.d2g_free(pi); self.L.d2g_free(pv)
        return indptr, idx, val

    def set_filterset(self, seq: np.ndarray, rec_off: np.ndarray, p: SketchParams) -> int:
        """--filterset from the records of a FASTX file (d2g_set_filterset); returns the number of hashed values kept."""
        seq = np.ascontiguousarray(seq, dtype=np.uint8); rec_off = np.ascontiguousarray(rec_off, dtype=np.uint64)
        n = u64_t(0)
        _check(self.L.d2g_set_filterset(self.h, C.byref(p), _ptr(seq), _ptr(rec_off), len(rec_off) - 1, C.byref(n)))
        return int(n.value)

    def set_filterset_values(self, values: np.ndarray):
        values = np.ascontiguousarray(values, dtype=np.uint64)
        _check(self.L.d2g_set_filterset_values(self.h, _ptr(values), len(values)))

    def clear_filterset(self):
        _check(self.L.d2g_clear_filterset(self.h))

    def cmp_rows_sharded(self, p: CmpParams, local_regs: np.ndarray, local_cards: np.ndarray, local_begin: int, row_begin: int, row_end: int, out: np.ndarray):
        """Collective: this rank's block of host registers in, rows [row_begin, row_end) of the whole matrix out (d2g_cmp_rows_sharded)."""
        _check(self.L.d2g_cmp_rows_sharded(self.h, C.byref(p), _ptr(local_regs), _ptr(local_cards), local_begin, len(local_cards), row_begin, row_end, _ptr(out)))
        return out
