"""CPU: the pipeline of the reference-side driver (rnamotif_b200/host/rm_gpu_main.c)
against the reference binary on inputs FN_fgetseq treats specially
(src/dbutil.c:42-128): several files, a file that does not begin with '>', an entry
without a name, records longer than -maxslen, headers without a definition, '>' in
the middle of a line, no newline at the end.  Where the device reader would not
read like FN_fgetseq the driver hands over to the host reader at that batch; stdout
must be the reference's byte for byte either way.  (rnamotif_hostcheck = the driver
linked against the oracle-backed stand-in of the C ABI, oracle/gm_mock.c.)"""
import os
import subprocess

import numpy as np
import pytest

import helpers
from rnamotif_b200 import synth

BIN = os.path.join(helpers.REF, "rnamotif_hostcheck")
REF_BIN = os.path.join(helpers.REF, "rnamotif")
DATA = os.path.join(helpers.REF, "data")
have = os.path.exists(BIN) and os.path.exists(REF_BIN)
pytestmark = pytest.mark.skipif(not have, reason="oracle/_ref (rnamotif, rnamotif_hostcheck) not built")


def fasta(seed, lengths, width=60):
    ids, seq, off = synth.random_records(seed, lengths, planted=True, iupac_rate=0.002)
    out = []
    for i, sid in enumerate(ids):
        out.append(b">" + sid.encode() + b" record %d of seed %d\n" % (i, seed))
        s = seq[off[i]:off[i + 1]].tobytes()
        out += [s[k:k + width] + b"\n" for k in range(0, len(s), width)]
    return b"".join(out)


def both(tmp_path, files, descr="trna", flags=(), env=None):
    paths = []
    for i, data in enumerate(files):
        p = tmp_path / ("in%d.fastn" % i)
        p.write_bytes(data)
        paths.append(str(p))
    e = dict(os.environ, EFNDATA=os.path.join(DATA, "efndata"))
    cmd = [*flags, "-descr", os.path.join(DATA, "test", descr + ".descr"), *paths]
    ref = subprocess.run([REF_BIN, *cmd], env=e, capture_output=True, timeout=600)
    e.update(env or {})
    got = subprocess.run([BIN, *cmd], env=e, capture_output=True, timeout=600)
    return ref, got


@pytest.mark.parametrize("batch", ["128000000", "20000"])
def test_several_files_and_small_batches(tmp_path, batch):
    files = [fasta(1, [9000, 300, 0, 15000]), fasta(2, [40000]), fasta(3, [5, 7000, 7000, 64])]
    ref, got = both(tmp_path, files, env={"GPUMOTIF_BATCH_NT": batch, "GPUMOTIF_DEVICES": "0,0"})
    assert got.returncode == ref.returncode == 0
    assert len(ref.stdout) > 0 and got.stdout == ref.stdout


@pytest.mark.parametrize("reader", ["device", "host"])
def test_odd_fasta_layouts(tmp_path, reader):
    a = fasta(4, [12000, 8000])
    odd = (b">first   \n" + a.split(b"\n", 1)[1]                                   # no definition, trailing blanks
           + b">mid\tdef with\ttabs\nacgu" + b"ACGUNRY\n>inline acgt\n" + fasta(5, [9000]).split(b"\n", 1)[1]
           + b">last no newline at the end\n" + fasta(6, [6000]).split(b"\n", 1)[1].rstrip(b"\n"))
    ref, got = both(tmp_path, [odd], env={"GPUMOTIF_BATCH_NT": "15000", **({"GPUMOTIF_READER": "host"} if reader == "host" else {})})
    assert got.returncode == ref.returncode == 0
    assert len(ref.stdout) > 0 and got.stdout == ref.stdout


def test_handover_to_the_host_reader(tmp_path):
    """An unnamed entry ends its file for FN_fgetseq (:62-66), a file that does not
    begin with '>' is skipped (:56-60), and a record longer than -maxslen is truncated
    (:104-125): the pipeline stops at the batch where it meets one of these and the
    host reader -- the reference's own -- carries on from that batch's offset."""
    good = fasta(7, [9000, 9000, 9000])
    unnamed = fasta(8, [9000, 5000]) + b">\nacgtacgtacgt\n" + fasta(9, [9000])
    nohdr = b"acgtacgt\n" + fasta(10, [4000])
    for files, flags in (([good, unnamed, good], ()), ([nohdr, good], ()), ([good, good], ("-N", "8000"))):
        ref, got = both(tmp_path, files, flags=flags, env={"GPUMOTIF_BATCH_NT": "12000"})
        assert got.returncode == ref.returncode
        assert got.stdout == ref.stdout
        assert len(ref.stdout) > 0
