"""Host-packed upload (gm_db_upload_chars_hostpack / gm_host_pack, include/gpumotif.h).

CPU: the packer against a numpy restatement of the device's code table
(rnamotif_b200/csrc/gm_kernel.cuh code_of_char: case folded, u = t, other characters 0)
on every byte value, ragged lengths, unaligned buffers, with and without AVX2, one and
several threads.  GPU: the candidate stream after a host-packed upload equals the
oracle's and the character upload's, chunk-streamed and not."""
import os
import subprocess
import sys

import numpy as np
import pytest

from rnamotif_b200 import gpumotif, synth
from oracle import oracle_port
import helpers

CODES = {"a": 1, "c": 2, "g": 4, "t": 8, "u": 8, "r": 5, "y": 10, "m": 3, "k": 12, "s": 6, "w": 9, "h": 11, "b": 14,
         "v": 7, "d": 13, "n": 15}


def ref_pack(seq):
    lut = np.zeros(256, dtype=np.uint8)
    for ch, code in CODES.items():
        lut[ord(ch)] = lut[ord(ch.upper())] = code
    c = lut[seq]
    if c.size & 1:
        c = np.concatenate([c, np.zeros(1, dtype=np.uint8)])
    return (c[0::2] | (c[1::2] << 4)).astype(np.uint8)


def corpus(n, seed):
    rng = np.random.default_rng(seed)
    seq = np.frombuffer(b"acgtACGUnNryRYmkswhbvdxeXZ*- 09@[`{", dtype=np.uint8)[rng.integers(0, 35, size=n)].copy()
    k = min(n, 256)
    seq[:k] = np.arange(256, dtype=np.uint8)[:k]  # every byte value once
    return seq


@pytest.mark.parametrize("n", [0, 1, 2, 63, 64, 65, 127, 4097, 70001, (1 << 20) + 77])
@pytest.mark.parametrize("threads", [1, 3, 8])
def test_host_pack_matches_code_table(n, threads):
    seq = corpus(n, n + threads)
    assert (gpumotif.host_pack(seq, threads) == ref_pack(seq)).all()


def test_host_pack_unaligned_source_and_scalar_path():
    seq = corpus(300000, 9)
    for off in (1, 7, 33):
        assert (gpumotif.host_pack(seq[off:], 2) == ref_pack(seq[off:])).all()
    # the table loop (CPUs without AVX2) gives the same bytes
    code = ("import sys, numpy as np; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
            "from rnamotif_b200 import gpumotif; import test_hostpack as t\n"
            "s = t.corpus(200001, 4); assert (gpumotif.host_pack(s, 2) == t.ref_pack(s)).all()\n"
            % (helpers.ROOT, helpers.HERE))
    env = dict(os.environ, GPUMOTIF_NO_AVX2="1")
    subprocess.run([sys.executable, "-c", code], check=True, env=env)


@pytest.mark.gpu
@pytest.mark.parametrize("chunk", ["default", "16384"])
@pytest.mark.parametrize("name", ["trna", "score.1", "pk1", "qu+tr"])
def test_hostpacked_upload_matches_oracle(name, chunk, monkeypatch):
    if chunk != "default":
        monkeypatch.setenv("GPUMOTIF_CHUNK_NT", chunk)  # several chunks: the scan takes them as they are published
    monkeypatch.setenv("GPUMOTIF_PACK_THREADS", "3")
    plan = helpers.load_plan(name)
    rng = np.random.default_rng(11)
    lengths = list(rng.integers(0, 6000, size=40)) + [150001, 3, 70000]
    ids, seq, off = synth.random_records(23, lengths, planted=True, iupac_rate=0.002)
    both = bool(gpumotif.plan_field(plan, 8))
    ref, _ = oracle_port.scan_db(plan, seq, off, both)
    ms = gpumotif.MotifSearch(plan)
    by_chars = ms.find_motif(seq, off)
    ms.upload_hostpack(seq, off)
    first = ms.scan()
    again = ms.scan()                      # same upload, plain path
    with pytest.raises(gpumotif.GpuMotifError):
        ms.hit_windows(0, 0)               # the device holds no characters after this upload
    ms.upload_hostpack(seq[:int(off[5])], off[:6])   # a smaller batch over the old one
    small = ms.scan()
    back = ms.find_motif(seq, off)         # and the character upload again
    ms.close()
    helpers.assert_same_hits(by_chars, ref, f"{name}: character upload")
    helpers.assert_same_hits(first, ref, f"{name}: host-packed upload")
    helpers.assert_same_hits(again, ref, f"{name}: host-packed upload, rescan")
    helpers.assert_same_hits(back, ref, f"{name}: character upload after a host-packed one")
    ref_small, _ = oracle_port.scan_db(plan, seq[:int(off[5])], off[:6], both)
    helpers.assert_same_hits(small, ref_small, f"{name}: smaller host-packed batch")


@pytest.mark.gpu
def test_hostpacked_upload_never_scanned_then_destroyed():
    """An upload nobody scans: destroy (and a following upload) join the uploader thread."""
    plan = helpers.load_plan("trna")
    ids, seq, off = synth.random_records(3, [200000, 50], planted=True)
    ms = gpumotif.MotifSearch(plan)
    ms.upload_hostpack(seq, off)
    ms.upload_hostpack(seq, off)
    ms.close()


@pytest.mark.gpu
@pytest.mark.parametrize("frac", ["0.3", "0.6", "1.0"])
def test_hostpacked_upload_from_pinned_memory_splits_chunks(frac, monkeypatch):
    """From pinned memory a share of every chunk is packed on the host and the rest goes over as
    characters (packed on the device) at the same time; the candidate stream does not depend on the share."""
    import torch
    monkeypatch.setenv("GPUMOTIF_CHUNK_NT", "16384")
    monkeypatch.setenv("GPUMOTIF_PACK_THREADS", "4")
    monkeypatch.setenv("GPUMOTIF_PACK_FRAC", frac)
    plan = helpers.load_plan("trna")
    rng = np.random.default_rng(12)
    lengths = list(rng.integers(0, 6000, size=30)) + [150001, 7, 90000]
    ids, seq, off = synth.random_records(29, lengths, planted=True, iupac_rate=0.002)
    ref, _ = oracle_port.scan_db(plan, seq, off, True)
    pinned = torch.empty(len(seq), dtype=torch.uint8, pin_memory=True)
    pinned.numpy()[:] = seq
    ms = gpumotif.MotifSearch(plan)
    for _ in range(2):
        ms.upload_ptr(pinned.data_ptr(), off, host_pack=True)
        hits = ms.scan()
        helpers.assert_same_hits(hits, ref, f"pinned host-packed upload, share {frac}")
    st = ms.stats()
    ms.close()
    n = int(off[-1])
    if frac == "1.0":
        assert st.h2d_bytes < 0.55 * n + 8 * len(off) + 64
    else:
        assert 0.5 * n < st.h2d_bytes < n
