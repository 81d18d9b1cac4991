"""CPU-side checks of the drop-in boundary: libd2gpu.so loads, exports every symbol include/d2gpu.h
declares, and refuses to run without a CUDA device (no CPU fallback)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def test_library_exports_every_declared_symbol():
    from dashing2_b200 import capi
    hdr = open(os.path.join(ROOT, "include", "d2gpu.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(d2g_[a-z0-9_]+)\s*\(", hdr)) - {"d2g_sink_fn"}
    assert declared == set(capi.EXPORTS), declared ^ set(capi.EXPORTS)
    if not os.path.exists(capi.LIB_PATH):
        pytest.skip("libd2gpu.so not built (run __graft_entry__.build())")
    lib = ctypes.CDLL(capi.LIB_PATH)
    for s in declared:
        assert hasattr(lib, s), s


def test_no_cpu_fallback_without_device():
    from dashing2_b200 import capi
    if not os.path.exists(capi.LIB_PATH):
        pytest.skip("libd2gpu.so not built")
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    with pytest.raises(capi.D2GError, match="no CUDA device|CPU fallback"):
        capi.Context(0)


def test_host_only_entry_points_work_without_device():
    """d2g_opmh_finalize / d2g_count_kmers / d2g_cmp_output_size are host arithmetic, callable on CPU."""
    import numpy as np
    from dashing2_b200 import capi
    if not os.path.exists(capi.LIB_PATH):
        pytest.skip("libd2gpu.so not built")
    import oracle_lib as O
    L = capi.load()
    rng = np.random.default_rng(0)
    regs = rng.integers(0, 2**63, size=(3, 64), dtype=np.uint64) * 2
    regs[0, :5] = np.uint64(2**64 - 1); regs[1, 7] = 0
    sig = np.empty((3, 64)); card = np.empty(3)
    assert L.d2g_opmh_finalize(regs.ctypes.data, 3, 64, sig.ctypes.data, card.ctypes.data) == 0
    Lo = O.lib()
    for i in range(3):
        s = np.empty(64); Lo.d2o_opmh_sigs(regs[i].copy(), 64, s)
        assert np.array_equal(s, sig[i]) and card[i] == Lo.d2o_opmh_card(regs[i].copy(), 64)
    p = capi.CmpParams(64, 0, 0, 31, 0, 10, 0)
    assert L.d2g_cmp_output_size(ctypes.byref(p)) == 45
    off = np.array([0, 10, 40, 100], dtype=np.uint64)
    assert L.d2g_count_kmers(off.ctypes.data, 3, 31) == 0 + 0 + 30


def test_front_end_fails_loudly_without_device(tmp_path):
    """dashing2-gpu has no CPU path either: without a CUDA device it exits non-zero with the library's message (and writes no output)."""
    import subprocess
    exe = os.path.join(ROOT, "dashing2_b200", "bin", "dashing2-gpu")
    if not os.path.exists(exe):
        pytest.skip("front-end not built")
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    fa = tmp_path / "x.fa"; fa.write_text(">r\n" + "ACGT" * 50 + "\n")
    out = tmp_path / "x.stk"
    r = subprocess.run([exe, "sketch", "-k31", "-S64", "-o", str(out), str(fa)], capture_output=True, text=True)
    assert r.returncode != 0 and ("no CUDA device" in r.stderr or "CPU fallback" in r.stderr), r.stderr
    assert not out.exists()
    # option errors are reported before any device work
    r = subprocess.run([exe, "sketch", "--entmin", str(fa)], capture_output=True, text=True)
    assert r.returncode != 0 and "not supported" in r.stderr


def _pack_reference(data: bytes):
    """numpy restatement of the packed layout (include/d2gpu.h): codes u64 words, first base in bits 63:62; mask u32, first base in bit 31."""
    import numpy as np
    from dashing2_b200 import capi
    L = capi.load()
    n = len(data)
    nw = int(L.d2g_packed_words(n))
    b = np.zeros(nw * 32, dtype=np.uint8)
    b[:n] = np.frombuffer(data, dtype=np.uint8)
    x = (b >> 1) & 3
    code = (x ^ (x >> 1)).astype(np.uint64)
    u = b & 0xDF
    inv = ~((u == ord("A")) | (u == ord("C")) | (u == ord("G")) | (u == ord("T")))
    sh = (np.uint64(62) - np.uint64(2) * np.arange(32, dtype=np.uint64))
    codes = (code.reshape(nw, 32) << sh).sum(axis=1, dtype=np.uint64)
    msh = (np.uint32(31) - np.arange(32, dtype=np.uint32))
    mask = (inv.reshape(nw, 32).astype(np.uint32) << msh).sum(axis=1, dtype=np.uint32)
    return codes, mask


@pytest.mark.parametrize("isa", ["0", "1", "2"])
def test_host_packer_matches_layout(isa):
    """d2g_pack_sequences (scalar / AVX2 / AVX-512 paths, pieces of every alignment) against the numpy restatement."""
    import subprocess, sys, textwrap
    from dashing2_b200 import capi
    if not os.path.exists(capi.LIB_PATH):
        pytest.skip("libd2gpu.so not built")
    flags = open("/proc/cpuinfo").read()
    if (isa == "2" and "avx512bw" not in flags) or (isa == "1" and "avx2" not in flags):
        pytest.skip("ISA not available on this CPU")
    # the ISA is latched on first use: run in a child with D2G_PACK_ISA set
    code = textwrap.dedent("""
        import sys, numpy as np
        sys.path[:0] = [%r, %r]
        from dashing2_b200 import capi
        from test_cabi_exports import _pack_reference
        rng = np.random.default_rng(7)
        for trial in range(6):
            n_pieces = int(rng.integers(1, 40))
            pieces = []
            for i in range(n_pieces):
                ln = int(rng.choice([0, 1, 5, 31, 32, 33, 63, 64, 65, 127, 1000, 4097, 70000, 1 << 21]))
                alphabet = np.frombuffer(b"ACGTacgtNnRYKM-*\\x00", dtype=np.uint8) if (trial %% 2) else np.frombuffer(b"ACGT", dtype=np.uint8)
                pr = np.full(len(alphabet), 0.02); pr[:4] = 1.0; pr /= pr.sum()
                pieces.append(alphabet[rng.choice(len(alphabet), size=ln, p=pr)].tobytes())
            codes, mask, nz = capi.pack_sequences(pieces)
            data = b"".join(pieces)
            ec, em = _pack_reference(data)
            nreal = (len(data) + 31) // 32
            assert np.array_equal(codes[:nreal], ec[:nreal]), trial
            assert np.array_equal(mask[:nreal], em[:nreal]), trial
            assert (mask[nreal:] == 0xFFFFFFFF).all() and (codes[nreal:] == 0).all()
            inv = np.frombuffer(data, dtype=np.uint8) & 0xDF
            bad = ~np.isin(inv, np.frombuffer(b"ACGT", dtype=np.uint8))
            words_bad = len(np.unique(np.nonzero(bad)[0] // 32))
            assert nz == words_bad, (nz, words_bad)
        print("ok")
    """) % (ROOT, os.path.join(ROOT, "tests"))
    env = dict(os.environ, D2G_PACK_ISA=isa, D2G_HOST_THREADS="4")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env)
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-2000:]
