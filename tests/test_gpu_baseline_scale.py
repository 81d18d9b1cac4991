"""Parity at the BASELINE configs' own scale (BASELINE.json configs[0..4]): the paths that only large inputs reach -- the guessed
Full SetSketch bound with threshold sharing between CTAs, the fast windowed kernel over thousands of tiles, counting sketches over
tens of millions of k-mers, comparison jobs with more sketches than 16-bit ranks, LSH graphs over 70 000 sketches -- against the
unmodified reference binary (oracle/_ref, travels with the snapshot) where it finishes in seconds, else against the oracle."""
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle_lib as O
from conftest import ROOT
from gpu_util import ctx, pack_batch

pytestmark = pytest.mark.gpu
EXE = os.path.join(ROOT, "dashing2_b200", "bin", "dashing2-gpu")


def u64(a):
    return np.ascontiguousarray(a).view(np.uint64)


def _ref():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import refbin
    if refbin.ref_binary() is None:
        pytest.skip("reference binary not available on this host")
    return refbin


def _run(exe_args, env=None):
    r = subprocess.run(exe_args, capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    return r


def test_config1_cli_outputs_byte_identical_to_reference_binary(tmp_path):
    """configs[0]: `sketch -k31 -S1024 --cmpout` over 64 genomes x 1 Mbp (seed 12345): PHYLIP text, binary matrix and the stacked
    -o file of dashing2-gpu against the same argv run through the reference binary, byte for byte."""
    from dashing2_b200 import synth
    refbin = _ref()
    paths = synth.write_fasta_set(str(tmp_path / "fa"), 64, 1_000_000, seed=12345, n_families=1)
    flist = tmp_path / "files.txt"; flist.write_text("\n".join(paths) + "\n")
    outs = {}
    for tag in ("ref", "gpu"):
        for kind, extra in (("phylip", ["--phylip"]), ("bin", ["--binary-output"])):
            stk = str(tmp_path / f"{tag}_{kind}.stk"); mat = str(tmp_path / f"{tag}_{kind}.out")
            argv = ["sketch", "-k31", "-S1024", "-p8", "-F", str(flist), "-o", stk, "--cmpout", mat] + extra
            if tag == "ref":
                refbin.run_ref(argv, threads=8)
            else:
                _run([EXE] + argv)
            outs[(tag, kind)] = (open(stk, "rb").read(), open(mat, "rb").read(), open(stk + ".names.txt", "rb").read())
    for kind in ("phylip", "bin"):
        r, g = outs[("ref", kind)], outs[("gpu", kind)]
        assert g[1] == r[1], f"{kind}: distance output differs"
        assert g[0] == r[0], f"{kind}: stacked sketch file differs"
        assert g[2] == r[2], f"{kind}: names file differs"


def test_config2_full_setsketch_registers_equal_reference_binary(tmp_path):
    """configs[1] per genome: 4 genomes x 5 Mbp, -k31 -w51 --full-setsketch -S4096.  Registers bit-equal to the reference binary; the
    library reports that no entity took the boot pass (the guessed bound held and was verified)."""
    from dashing2_b200 import synth
    refbin = _ref()
    paths = synth.write_fasta_set(str(tmp_path / "fa"), 4, 5_000_000, seed=2, n_families=2)
    flist = tmp_path / "files.txt"; flist.write_text("\n".join(paths) + "\n")
    argv = ["sketch", "-k31", "-w51", "--full-setsketch", "-S4096", "-p4", "-F", str(flist)]
    refbin.run_ref(argv + ["-o", str(tmp_path / "ref.ss")], threads=4)
    r = _run([EXE] + argv + ["-o", str(tmp_path / "gpu.ss")], env=dict(os.environ, D2G_DEBUG="1"))
    assert "0 of 4 entities take the boot pass" in r.stderr, r.stderr[-1500:]
    a = np.fromfile(tmp_path / "ref.ss", dtype=np.uint64); b = np.fromfile(tmp_path / "gpu.ss", dtype=np.uint64)
    assert a[0] == b[0] == 4 and a[1] == b[1] == 4096
    assert np.array_equal(a[2 + 4:], b[2 + 4:]), "registers differ from the reference binary"
    np.testing.assert_allclose(b[2:6].view(np.float64), a[2:6].view(np.float64), rtol=1e-12)


@pytest.mark.parametrize("mode", ["bmh", "pmh"])
def test_config3_counting_sketches_20mbp_match_oracle(mode):
    """configs[2] per genome: 20 Mbp with a tenth of it duplicated (k-mer multiplicities > 1), S = 8192, exact counting."""
    from dashing2_b200 import synth
    S, k = 8192, 31
    gen = [[s.tobytes()] for _, s in synth.family_genomes(2, 20_000_000, seed=3, dup_frac=0.1)]
    c = ctx()
    seq, off, ent = pack_batch(gen)
    r = c.sketch_batch(seq, off, ent, len(gen), c.params(mode=mode, S=S, k=k))
    for e in range(len(gen)):
        o = O.weighted_sketch(O.hash_stream(gen[e][0], k), mode, S)
        assert np.array_equal(u64(r["sig"][e]), u64(o["sig"])), (mode, e)
        assert r["card"][e] == o["card"]


def test_config4_shape_more_sketches_than_ranks_rows_match_oracle():
    """configs[3] shape: S = 1024 and more sketches (72 000) than one comparison job holds, so rows are computed by block-pair jobs
    whose 16-bit codes derive from one global ranking.  500 sampled rows of the symmetric matrix against the oracle's compare()."""
    from dashing2_b200 import synth
    n, S = 72_000, 1024
    regs, cards = synth.synthetic_sketches(n, S, seed=4, n_families=700)
    cards = cards * (1 + (np.arange(n) % 5))
    c = ctx()
    p = c.cmp_params(S, n, "symmetric", "containment", k=31)
    L = O.lib()
    buf = np.empty((n + 1, S)); buf[1:] = regs
    cbuf = np.empty(n + 1); cbuf[1:] = cards
    out = np.empty(n, dtype=np.float32)
    for r0 in (0, 31_500, 63_100, 71_700):
        rows = c.cmp_rows(regs, cards, p, r0, r0 + 125)
        at = 0
        for i in range(r0, r0 + 125):
            buf[0] = regs[i]; cbuf[0] = cards[i]
            L.d2o_panel(buf, cbuf, 1, n, S, O.MEASURES["containment"], 31, 0, out)
            m = n - i - 1
            assert np.array_equal(rows[at:at + m].view(np.uint32), out[i + 1:].view(np.uint32)), i
            at += m
        assert at == len(rows)


def test_config5_shape_topk_graph_70000_matches_oracle():
    """configs[4] shape: --topk 32 over 70 000 sketches, S = 1024 (index build, ordered candidate scan, bounded lists, refinement):
    the whole CSR against the oracle's sequential graph."""
    from dashing2_b200 import synth
    n, S, K = 70_000, 1024, 32
    regs, cards = synth.synthetic_sketches(n, S, seed=5, n_families=70)
    ep, ei, ev = O.topk(regs, cards, K, "similarity", k=31)
    gp, gi, gv = ctx().lsh_topk(regs, cards, K, "similarity", k=31)
    assert np.array_equal(gp, ep) and np.array_equal(gi, ei) and np.array_equal(gv.view(np.uint32), ev.view(np.uint32))
