"""GPU parity for d2g_lsh_graph: similarity-threshold neighbour graphs and --topk over compressed registers, against the CSR files of the
reference binary (-p1) and the oracle on seeded sketches."""
import os

import numpy as np
import pytest

import oracle_lib as O
from conftest import expected, GOLD
from gpu_util import ctx

pytestmark = pytest.mark.gpu


def same_csr(a, b):
    return np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2].view(np.uint32), b[2].view(np.uint32))


@pytest.mark.parametrize("tag,thr,measure", [("t0.5", 0.5, "similarity"), ("t0.8", 0.8, "similarity"), ("t0.3_containment", 0.3, "containment")])
def test_similarity_threshold_graph_matches_reference_golden(tag, thr, measure):
    """--similarity-threshold x (src/index_build.cpp:53-165 with topk = -1, src/refine.cpp:43-68): the reference binary's CSR, byte for byte."""
    z = np.load(os.path.join(GOLD, "inputs", "sk600x64.npz"))
    exp = O.read_csr(expected(f"nnthr_{tag}_sk600.csr"))
    got = ctx().lsh_graph(z["regs"], z["cards"], -1, thr, measure, k=32)
    assert same_csr(got, exp)
    # row ranges (how several GPUs share one graph) concatenate to it
    parts = [ctx().lsh_graph(z["regs"], z["cards"], -1, thr, measure, k=32, rows=r) for r in ((0, 250), (250, 600))]
    idx = np.concatenate([p[1] for p in parts]); val = np.concatenate([p[2] for p in parts])
    assert np.array_equal(idx, exp[1]) and np.array_equal(val.view(np.uint32), exp[2].view(np.uint32))


@pytest.mark.parametrize("n,S,thr,measure,cmp_kind", [(900, 64, 0.3, "similarity", 0), (1500, 128, 0.15, "similarity", 0), (700, 32, 0.5, "poisson_llr", 0),
                                                      (800, 64, 0.9, "symmetric_containment", 0), (600, 64, 0.2, "similarity", 1), (300, 16, 0.0, "similarity", 0)])
def test_similarity_threshold_graph_matches_oracle_seeded(n, S, thr, measure, cmp_kind):
    """Uncapped candidate lists (several hundred entries in the dense families), the 20-consecutive-failures cut, distances (v < x)."""
    from dashing2_b200 import synth
    regs, cards = synth.synthetic_sketches(n, S, seed=n + S, n_families=4, p_lo=0.02, p_hi=0.6)
    exp = O.nn_threshold(regs, cards, thr, measure, k=31, cmp_kind=cmp_kind)
    got = ctx().lsh_graph(regs, cards, -1, thr, measure, k=31, cmp_kind=cmp_kind)
    assert same_csr(got, exp)
    assert len(exp[1]) > 0


@pytest.mark.parametrize("tag,fd,bbit", [("fd1", 1, False), ("fd2_bbit", 2, True)])
def test_topk_with_fastcmp_matches_reference_golden(tag, fd, bbit):
    """--topk 8 --fastcmp N [--bbit-sigs]: index over the f64 signatures, refinement through the compressed compare branch
    (src/cmp_core.cpp:741-799,362-449)."""
    z = np.load(os.path.join(GOLD, "inputs", "sk600x64.npz"))
    c = ctx()
    creg, kind, a, b = c.make_compressed(z["regs"], fd, bbit)
    exp = O.read_csr(expected(f"topk8_{tag}_sk600.csr"))
    got = c.lsh_graph(creg, z["cards"], 8, 0., "similarity", k=32, cmp_kind=kind, index_regs=z["regs"], regbytes=float(fd), compressed_b=b)
    assert same_csr(got, exp)


@pytest.mark.parametrize("fd,bbit,measure,K", [(1, False, "similarity", 12), (2, False, "containment", 5), (4, True, "similarity", 20), (1, True, "poisson_llr", 7)])
def test_topk_with_fastcmp_matches_oracle_seeded(fd, bbit, measure, K):
    from dashing2_b200 import synth
    regs, cards = synth.synthetic_sketches(1000, 64, seed=77 + fd, n_families=5, p_lo=0.02, p_hi=0.5)
    c = ctx()
    creg, kind, a, b = c.make_compressed(regs, fd, bbit)
    ocreg, otrunc, oa, ob = O.make_compressed(regs, fd, bbit)
    assert np.array_equal(creg, ocreg)
    exp = O.topk_compressed(regs, ocreg, cards, K, fd, kind == 3, ob, measure, k=31)
    got = c.lsh_graph(creg, cards, K, 0., measure, k=31, cmp_kind=kind, index_regs=regs, regbytes=float(fd), compressed_b=b)
    assert same_csr(got, exp)


def test_graph_limits_fail_loudly():
    from dashing2_b200.capi import D2GError
    from dashing2_b200 import synth
    regs, cards = synth.synthetic_sketches(200, 16, seed=3, n_families=2)
    c = ctx()
    creg, kind, a, b = c.make_compressed(regs, 1, False)
    with pytest.raises(D2GError):       # compressed registers without the signatures the index is built over
        c.lsh_graph(creg, cards, 5, 0., "similarity", cmp_kind=2, regbytes=1.0, compressed_b=b)
    with pytest.raises(D2GError):
        c.lsh_topk(regs, cards, 0)
