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
    r = subprocess.run([exe, "sketch", "--similarity-threshold", "0.5", str(fa)], capture_output=True, text=True)
    assert r.returncode != 0 and "not supported" in r.stderr
