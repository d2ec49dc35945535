"""CPU: sharding math, and the N>1 path end to end on the `gloo` backend
(world_size 2): each rank scans its shard with the oracle port, rank 0 merges,
and the result equals the single-process scan."""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import oracle_port
from rnamotif_b200 import synth
import shard
import helpers


def test_shard_ranges_cover_and_balance():
    for total in (0, 1, 7, 1000, 12345678901):
        for world in (1, 2, 3, 8):
            r = shard.shard_ranges(total, world)
            assert r[0][0] == 0 and r[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(r[:-1], r[1:]))
            sizes = [hi - lo for lo, hi in r]
            assert max(sizes) - min(sizes) <= 1


def test_shard_records_and_merge_equal_single_scan():
    plan = helpers.load_plan("score.1")
    ids, seq, off = synth.random_records(3, [3000, 10, 0, 8000, 500, 12000, 64], planted=True)
    whole, _ = oracle_port.scan_db(plan, seq, off, True)
    parts, bases = [], []
    for a, b in shard.shard_records(off, 3):
        sub_off = off[a:b + 1] - off[a]
        sub_seq = seq[off[a]:off[b]]
        h, _ = oracle_port.scan_db(plan, sub_seq, sub_off, True)
        parts.append(h)
        bases.append(a)
    merged = shard.merge_hits(parts, bases)
    assert merged is not None and len(whole) > 0
    helpers.assert_same_hits(merged, whole, "record shards")


WORKER = r"""
import os, sys, pickle
import numpy as np
import torch.distributed as dist
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
from oracle import oracle_port
from rnamotif_b200 import synth
import shard
import helpers
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
plan = helpers.load_plan("trna")
ids, seq, off = synth.random_records(11, [20000, 300, 0, 15000, 9000, 64, 30000], planted=True)
a, b = shard.shard_records(off, world)[rank]
h, _ = oracle_port.scan_db(plan, seq[off[a]:off[b]], off[a:b + 1] - off[a], True)
gathered = [None] * world
dist.all_gather_object(gathered, (a, h))
if rank == 0:
    merged = shard.merge_hits([g[1] for g in gathered], [g[0] for g in gathered])
    whole, _ = oracle_port.scan_db(plan, seq, off, True)
    helpers.assert_same_hits(merged, whole, "gloo world_size 2")
    print("OK", len(whole))
dist.barrier()
dist.destroy_process_group()
"""


def test_two_rank_gloo_scan_matches_single_process(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29631", OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29631", str(script), helpers.ROOT],
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "OK" in r.stdout
