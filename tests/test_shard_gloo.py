"""world_size-2 gloo test of the multi-GPU host logic: file sharding, the register all-gather and the
equal-area row blocks whose per-rank outputs concatenate to the single-process matrix.  The per-rank
compute here is the oracle (CPU stand-in for the kernel; the NCCL + kernel version is bench.py --gpus N)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, tmp):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle_lib as O
    from dashing2_b200 import shard, synth
    g, S = 37, 128
    regs, cards = synth.synthetic_sketches(g * world, S, seed=5, n_families=4)
    mine = slice(rank * g, (rank + 1) * g)
    all_sig, all_card = shard.gather_registers(dist, torch.from_numpy(regs[mine].copy()), torch.from_numpy(cards[mine].copy()), world)
    assert np.array_equal(all_sig.numpy(), regs) and np.array_equal(all_card.numpy(), cards)
    n = g * world
    b = shard.equal_area_rows(n, world)
    full = O.allpairs(regs, cards, "symmetric", "containment")
    tri = lambda i: i * n - i * (i + 1) // 2
    part = full[tri(b[rank]):tri(b[rank + 1])]
    # what this rank would compute: rows [b[rank], b[rank+1]) against all columns j > i
    rows = []
    for i in range(b[rank], b[rank + 1]):
        rows.append(O.allpairs(np.vstack([regs[i:i + 1], regs[i + 1:]]), np.concatenate([cards[i:i + 1], cards[i + 1:]]), "panel", "containment", nq=n - i - 1)
                    if i + 1 < n else np.empty(0, dtype=np.float32))
    got = np.concatenate(rows) if rows else np.empty(0, dtype=np.float32)
    assert np.array_equal(got.view(np.uint32), part.view(np.uint32))
    np.save(os.path.join(tmp, f"part{rank}.npy"), got)
    dist.barrier()
    if rank == 0:
        cat = np.concatenate([np.load(os.path.join(tmp, f"part{r}.npy")) for r in range(world)])
        assert np.array_equal(cat.view(np.uint32), full.view(np.uint32))
        sizes = [tri(b[r + 1]) - tri(b[r]) for r in range(world)]
        assert max(sizes) - min(sizes) <= n
    dist.destroy_process_group()


def test_world2_gloo(tmp_path):
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)


def test_partition_helpers():
    from dashing2_b200 import shard
    for n in (1, 2, 7, 100, 10000):
        for parts in (1, 2, 3, 8):
            b = shard.equal_area_rows(n, parts)
            assert b[0] == 0 and b[-1] == n and all(x <= y for x, y in zip(b, b[1:]))
    b = shard.equal_area_rows(10000, 8)
    tri = lambda i: i * 10000 - i * (i + 1) // 2
    areas = [tri(b[r + 1]) - tri(b[r]) for r in range(8)]
    assert max(areas) / min(areas) < 1.01
    sh = shard.shard_files([5, 1, 9, 3, 7], 2)
    assert sh == [[2, 0, 1], [4, 3]] and sorted(sum(sh, [])) == [0, 1, 2, 3, 4]
