"""Locate / run the reference binary built by oracle/Makefile.ref (TEST + BASELINE INFRASTRUCTURE ONLY).

The binary is built in the dev container from /root/reference (unmodified) and travels to the GPU
box inside oracle/_ref/ (git-ignored, not gpurun-ignored).  Nothing here reads /root/reference at
run time; `ensure_built()` is only called from __graft_entry__.build() / golden generation, and only
compiles when /root/reference is present.
"""
from __future__ import annotations

import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFDIR = os.path.join(ROOT, "oracle", "_ref")


def _cpu_flags():
    try:
        with open("/proc/cpuinfo") as f:
            for ln in f:
                if ln.startswith("flags"):
                    return set(ln.split(":", 1)[1].split())
    except OSError:
        pass
    return set()


def ref_binary():
    """Path of the best reference binary this CPU can execute, or None."""
    flags = _cpu_flags()
    v4 = {"avx512f", "avx512bw", "avx512dq", "avx512vl", "avx512cd", "avx512_vpopcntdq"}
    cands = []
    if v4 <= flags:
        cands.append("dashing2-v4")
    if {"avx2", "bmi2", "fma"} <= flags:
        cands.append("dashing2-v3")
    for c in cands:
        p = os.path.join(REFDIR, c)
        if os.path.isfile(p) and os.access(p, os.X_OK):
            return p
    return None


def ensure_built(jobs: int = 8) -> bool:
    """Build oracle/_ref from /root/reference when that tree exists (dev container only)."""
    if not os.path.isdir("/root/reference/src"):
        return ref_binary() is not None
    subprocess.check_call(["make", "-s", "-f", "oracle/Makefile.ref", f"-j{jobs}"], cwd=ROOT)
    return ref_binary() is not None


def run_ref(args, cwd=None, threads=None, check=True, timeout=None):
    exe = ref_binary()
    if exe is None:
        raise RuntimeError("reference binary not available (oracle/_ref missing or CPU lacks AVX2)")
    env = dict(os.environ)
    if threads is not None:
        env["OMP_NUM_THREADS"] = str(threads)
    return subprocess.run([exe] + list(args), cwd=cwd, env=env, check=check, timeout=timeout,
                          stdout=subprocess.PIPE, stderr=subprocess.PIPE)
