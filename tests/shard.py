"""TEST HELPER: sharding of a scan across ranks as the range form of
gm_scan allows it; the product's own multi-GPU path deals whole batches of records to
the GPUs (rnamotif_b200/host/rm_gpu_main.c) and bench.py gives every rank its own database.

Start positions are independent, so a database shards with no exchange step:
the concatenated records are cut into `world` contiguous ranges of start
positions of (almost) equal size; rank r scans [lo_r, hi_r) with
gm_scan(ctx, lo_r, hi_r, strands) -- the motif-span halo is read from the
neighbouring nucleotides by the kernel itself -- and the per-rank candidate
lists are merged by the enumeration key (rec, comp, szero, seq).  This replaces
the reference's MPI file farm (src/mrnamotif.c:105-192), whose unit is a whole
database file and whose output order is arrival order.
"""
from __future__ import annotations

import numpy as np


def shard_ranges(total_nt: int, world: int):
    """[(lo, hi)] * world, contiguous, covering [0, total_nt), sizes differ by <= 1."""
    if world < 1:
        raise ValueError("world must be >= 1")
    base, rem = divmod(int(total_nt), world)
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < rem else 0)
        out.append((lo, hi))
        lo = hi
    return out


def shard_records(rec_off, world: int):
    """Record-granular variant: [(first_rec, last_rec_exclusive)] balanced by
    nucleotide count (what a file farm would do)."""
    rec_off = np.asarray(rec_off, dtype=np.int64)
    n = len(rec_off) - 1
    cuts = [0]
    for r in range(1, world):
        target = rec_off[-1] * r // world
        cuts.append(int(np.searchsorted(rec_off, target, side="left")))
    cuts.append(n)
    cuts = [min(max(c, 0), n) for c in cuts]
    for i in range(1, len(cuts)):
        cuts[i] = max(cuts[i], cuts[i - 1])
    return list(zip(cuts[:-1], cuts[1:]))


def merge_hits(parts, rec_base=None):
    """Concatenate per-rank structured hit arrays and restore the reference's
    enumeration order.  rec_base[i] is added to the record numbers of part i
    (for record-granular shards that number their records from 0)."""
    parts = [p.copy() for p in parts]
    if rec_base is not None:
        for p, b in zip(parts, rec_base):
            p["rec"] += np.uint32(b)
    parts = [p for p in parts if len(p)]
    if not parts:
        return None
    allh = np.concatenate(parts)
    order = np.lexsort((allh["seq"], allh["szero"], allh["comp"], allh["rec"]))
    return allh[order]
