"""Builds and runs the host-side checks of the two arithmetic headers the kernels rely on:
xf80.h (software x87 long double) and devlog.cuh (glibc-identical log)."""
import os
import subprocess

import pytest

from conftest import ROOT


@pytest.mark.parametrize("src,n", [("xf80_check.cpp", "400000"), ("devlog_check.cpp", "1000000")])
def test_header_matches_host_arithmetic(src, n, tmp_path):
    exe = str(tmp_path / src.replace(".cpp", ""))
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-ffp-contract=off", "-x", "c++", "-o", exe,
                           os.path.join(ROOT, "tests", "cpp", src)])
    out = subprocess.run([exe, n], capture_output=True, text=True)
    assert out.returncode == 0 and "OK" in out.stdout, out.stdout[-2000:]


def test_make_compressed_host_matches_oracle():
    """d2g_make_compressed is host-side x87 arithmetic (no device needed): same quantised registers and fitted (a, b) as the
    oracle restatement of make_compressed (src/cmp_core.cpp:209-322), including the b-bit variant and degenerate inputs."""
    import ctypes as C
    import numpy as np
    import oracle_lib as O
    from dashing2_b200 import capi
    L = capi.load()
    rng = np.random.default_rng(3)
    regs = rng.random((37, 96)) * 1e-4
    regs[2, :10] = 0.0; regs[5] = regs[4]
    for fd in (1, 2, 4):
        for bbit in (False, True):
            out = np.empty_like(regs); a = C.c_longdouble(-1); b = C.c_longdouble(-1); used = C.c_int32(-1)
            rc = L.d2g_make_compressed(regs.ctypes.data, None, regs.shape[0], regs.shape[1], float(fd), int(bbit), C.byref(a), C.byref(b), out.ctypes.data, C.byref(used))
            assert rc == 0
            oreg, trunc, oa, ob = O.make_compressed(regs, fd, bbit)
            assert used.value == trunc == int(bbit) and np.array_equal(out, oreg)
            if not bbit:
                assert a.value == oa.value and b.value == ob.value
    # all registers equal -> b = 1, a = value: quantisation degenerates (log1p(0) = 0) but must not crash and must agree
    flat = np.full((3, 8), 0.25)
    out = np.empty_like(flat); a = C.c_longdouble(-1); b = C.c_longdouble(-1); used = C.c_int32(-1)
    assert L.d2g_make_compressed(flat.ctypes.data, None, 3, 8, 1.0, 0, C.byref(a), C.byref(b), out.ctypes.data, C.byref(used)) == 0
    oreg, trunc, oa, ob = O.make_compressed(flat, 1, False)
    assert used.value == trunc and np.array_equal(out, oreg)


def test_bench_entry_points_parse():
    """bench.py and its per-config module import and parse their arguments without a GPU (the driver calls them by contract)."""
    import subprocess, sys, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--help"], capture_output=True, text=True)
    assert r.returncode == 0 and "--config" in r.stdout and "--impl" in r.stdout
    sys.path.insert(0, root)
    import bench_configs
    assert callable(bench_configs.run)
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--config", "4"], capture_output=True, text=True)
    assert r.returncode != 0 and "CUDA device" in (r.stderr + r.stdout)      # fails loudly without a GPU, no CPU fallback
