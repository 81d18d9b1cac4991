"""Helpers for the -m gpu parity tests: batch packing and a shared libd2gpu context."""
import numpy as np

import oracle_lib as O


def pack_batch(files_records):
    """files_records: list (one per entity) of lists of bytes records -> (seq u8, rec_off u64, rec_entity u32)."""
    offs = [0]
    ents = []
    chunks = []
    for e, recs in enumerate(files_records):
        for r in recs:
            chunks.append(np.frombuffer(r, dtype=np.uint8))
            offs.append(offs[-1] + len(r))
            ents.append(e)
    seq = np.concatenate(chunks) if chunks else np.empty(0, dtype=np.uint8)
    return seq, np.asarray(offs, dtype=np.uint64), np.asarray(ents, dtype=np.uint32)


def pack_files(paths):
    return pack_batch([O.read_fastx(p) for p in paths])


_ctx = None


def ctx():
    global _ctx
    if _ctx is None:
        from dashing2_b200 import capi
        _ctx = capi.Context(0)
    return _ctx
