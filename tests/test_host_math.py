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
