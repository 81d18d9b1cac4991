"""Seeded synthetic inputs for the sketch / cmp hot paths.

Genomes: one random ACGT ancestor of length L per family; genome g is the ancestor with i.i.d.
substitutions at rate r_g = 0.001 * (g mod 64 + 1).  FASTA: one record, 80-column lines, header
``>g<idx>``.  Sketch matrices (configs 4 and 5) are synthesised directly as f64[n][S].

Everything is driven by ``numpy.random.default_rng(seed)`` so that the same seed gives the same
bytes here and on the GPU box (nothing in here reads /root/reference).
"""
from __future__ import annotations

import os
import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def family_genomes(n: int, length: int, seed: int, n_families: int = 1, dup_frac: float = 0.0):
    """Yield (idx, uint8 ASCII array) for ``n`` genomes of ``length`` bp.

    Genome g belongs to family ``g % n_families`` and carries substitutions at rate
    0.001 * (g % 64 + 1).  ``dup_frac`` > 0 appends a copy of the first ``dup_frac`` of the genome
    (so k-mer multiplicities > 1 exist; config 3).
    """
    rng = np.random.default_rng(seed)
    ancestors = [rng.integers(0, 4, size=length, dtype=np.uint8) for _ in range(n_families)]
    for g in range(n):
        anc = ancestors[g % n_families]
        rate = 0.001 * (g % 64 + 1)
        codes = anc.copy()
        nmut = rng.binomial(length, rate)
        if nmut:
            pos = rng.integers(0, length, size=nmut)
            # substitute by a *different* base
            codes[pos] = (codes[pos] + rng.integers(1, 4, size=nmut, dtype=np.uint8)) & 3
        seq = _ACGT[codes]
        if dup_frac > 0:
            seq = np.concatenate([seq, seq[: int(length * dup_frac)]])
        yield g, seq


def fasta_bytes(name: str, seq: np.ndarray, width: int = 80) -> bytes:
    """One-record FASTA with ``width``-column lines."""
    n = len(seq)
    nfull, rem = divmod(n, width)
    body = np.empty(n + nfull + (1 if rem else 0), dtype=np.uint8)
    if nfull:
        blk = body[: nfull * (width + 1)].reshape(nfull, width + 1)
        blk[:, :width] = seq[: nfull * width].reshape(nfull, width)
        blk[:, width] = 10
    if rem:
        body[nfull * (width + 1): -1] = seq[nfull * width:]
        body[-1] = 10
    return b">" + name.encode() + b"\n" + body.tobytes()


def write_fasta_set(outdir: str, n: int, length: int, seed: int, n_families: int = 1,
                    dup_frac: float = 0.0):
    """Write g<idx>.fa files; returns the list of paths (in index order)."""
    os.makedirs(outdir, exist_ok=True)
    paths = []
    for g, seq in family_genomes(n, length, seed, n_families, dup_frac):
        p = os.path.join(outdir, f"g{g}.fa")
        with open(p, "wb") as f:
            f.write(fasta_bytes(f"g{g}", seq))
        paths.append(p)
    return paths


def synthetic_sketches(n: int, sketchsize: int, seed: int, n_families: int = 1000,
                       p_lo: float = 0.05, p_hi: float = 0.95):
    """f64[n][S] register matrix + cardinalities (configs 4/5): family base row ``rng.random(S)``;
    each register resampled with probability p_g in [p_lo, p_hi]."""
    rng = np.random.default_rng(seed)
    nf = max(1, min(n_families, n))
    base = rng.random((nf, sketchsize))
    out = np.empty((n, sketchsize), dtype=np.float64)
    chunk = 4096
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        fam = np.arange(s, e) % nf
        rows = base[fam]
        p = rng.uniform(p_lo, p_hi, size=(e - s, 1))
        resample = rng.random((e - s, sketchsize)) < p
        fresh = rng.random((e - s, sketchsize))
        out[s:e] = np.where(resample, fresh, rows)
    cards = np.full(n, 1e6, dtype=np.float64)
    return out, cards


def write_stacked(path: str, regs: np.ndarray, cards: np.ndarray, names=None):
    """Reference stacked sketch file: u64 n, u64 S, f64 card[n], f64 reg[n][S]
    (/root/reference/src/sketch_core.cpp:130-139) + ``path.names.txt``."""
    n, s = regs.shape
    with open(path, "wb") as f:
        np.array([n, s], dtype=np.uint64).tofile(f)
        np.asarray(cards, dtype=np.float64).tofile(f)
        np.ascontiguousarray(regs, dtype=np.float64).tofile(f)
    if names is not None:
        with open(path + ".names.txt", "w") as f:
            f.write("#Name\tCardinality\n")
            for nm, c in zip(names, cards):
                f.write("%s\t%0.24g\n" % (nm, c))
