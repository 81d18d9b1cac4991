"""Host-side sharding of the two hot paths over one process per GPU.

sketch : input files are independent -> files sorted by descending size (get_filesizes,
         /root/reference/src/sketch_core.cpp:175-184) and dealt round-robin; no collective.
cmp    : one all-gather of the register matrix (+ cardinalities), then rank r computes a contiguous
         block of output rows; for the condensed triangle the blocks have equal *pair* counts.
"""
from __future__ import annotations

import numpy as np


def equal_area_rows(n: int, parts: int):
    """Row boundaries b[0..parts] giving each part (almost) the same number of upper-triangle pairs."""
    total = n * (n - 1) // 2
    b = [0]
    for r in range(1, parts):
        target = total * r // parts
        lo, hi = 0, n
        while lo < hi:
            mid = (lo + hi) // 2
            if mid * n - mid * (mid + 1) // 2 < target:
                lo = mid + 1
            else:
                hi = mid
        b.append(max(lo, b[-1]))
    b.append(n)
    return b


def equal_rows(n_rows: int, parts: int):
    """Panel / asymmetric: equal row counts."""
    return [n_rows * r // parts for r in range(parts + 1)]


def shard_files(sizes, world: int):
    """Largest-first round-robin assignment; returns per-rank lists of original file indices."""
    order = sorted(range(len(sizes)), key=lambda i: (-sizes[i], i))
    return [order[r::world] for r in range(world)]


def gather_registers(dist, sig, card, world: int):
    """All-gather [g][S] registers and [g] cards from every rank (equal g per rank) -> [world*g][S], [world*g]."""
    import torch
    if world == 1:
        return sig, card
    all_sig = torch.empty((sig.shape[0] * world, sig.shape[1]), dtype=sig.dtype, device=sig.device)
    all_card = torch.empty(card.shape[0] * world, dtype=card.dtype, device=card.device)
    dist.all_gather_into_tensor(all_sig, sig.contiguous())
    dist.all_gather_into_tensor(all_card, card.contiguous())
    return all_sig, all_card
