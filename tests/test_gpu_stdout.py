"""-m gpu: byte-identical rnamotif stdout.  rnamotif_gpu (reference front end +
score/efn/printer on the host, search in libgpumotif) against
  (a) the committed md5 of the reference binary's raw stdout for each of the
      24 `make test` command lines (tests/golden/make_test_md5.json; the same
      runs pass the reference's test/*.chk goldens, oracle/check_goldens.sh), and
  (b) the reference binary itself (oracle/_ref/rnamotif), run side by side.
Needs the prebuilt programs and test inputs under oracle/_ref/ and
rnamotif_b200/host/_build/ (they travel with the snapshot)."""
import hashlib
import json
import os
import subprocess

import pytest

import helpers

pytestmark = pytest.mark.gpu

GPU_BIN = os.path.join(helpers.ROOT, "rnamotif_b200", "host", "_build", "rnamotif_gpu")
REF_BIN = os.path.join(helpers.REF, "rnamotif")
DATA = os.path.join(helpers.REF, "data")
MD5 = json.load(open(os.path.join(helpers.GOLDEN, "make_test_md5.json")))

have = os.path.exists(GPU_BIN) and os.path.exists(os.path.join(DATA, "test", "gbrna.111.0.fastn"))


def run(binary, name, extra_env=None):
    flags = ["-sh", "-context", "-Dctx_maxlen=5"] if name.endswith(".strict") else []
    env = dict(os.environ, EFNDATA=os.path.join(DATA, "efndata"))
    env.update(extra_env or {})
    r = subprocess.run([binary, *flags, "-descr", name + ".descr", "gbrna.111.0.fastn"],
                       cwd=os.path.join(DATA, "test"), env=env, capture_output=True, timeout=1200)
    assert r.returncode == 0, r.stderr.decode(errors="replace")[-2000:]
    return r.stdout


@pytest.mark.skipif(not have, reason="oracle/_ref or rnamotif_gpu not built")
@pytest.mark.parametrize("name", sorted(MD5))
def test_make_test_stdout_md5(name):
    out = run(GPU_BIN, name)
    assert hashlib.md5(out).hexdigest() == MD5[name]["md5"], f"{name}: stdout differs from the reference"


@pytest.mark.skipif(not (have and os.path.exists(REF_BIN)), reason="reference binary not built")
@pytest.mark.parametrize("name", ["trna", "pk1", "qu+tr.strict", "getbest"])
def test_stdout_equals_reference_binary_small_batches(name):
    # small batches: records split across many uploads, score state carried across
    out = run(GPU_BIN, name, {"GPUMOTIF_BATCH_NT": "300000"})
    ref = run(REF_BIN, name)
    assert out == ref


@pytest.mark.skipif(not have, reason="oracle/_ref or rnamotif_gpu not built")
@pytest.mark.parametrize("name", ["trna", "getbest", "pk1.strict"])
def test_stdout_with_sharded_scan(name):
    """The host driver's in-process sharding (GPUMOTIF_DEVICES): the batch is cut
    into ranges of start positions, one context each, and the candidate lists are
    merged before the (stateful) score replay.  Two contexts on device 0 exercise
    the same code as two GPUs; with more than one GPU visible use them."""
    import torch
    devs = "0,1,0" if torch.cuda.device_count() > 1 else "0,0,0"
    out = run(GPU_BIN, name, {"GPUMOTIF_DEVICES": devs, "GPUMOTIF_BATCH_NT": "1000000"})
    assert hashlib.md5(out).hexdigest() == MD5[name]["md5"]


PRUNE_BIN = os.path.join(helpers.REF, "rmprune")


@pytest.mark.skipif(not (have and os.path.exists(REF_BIN) and os.path.exists(PRUNE_BIN)),
                    reason="reference binaries (rnamotif, rmprune) not built")
@pytest.mark.parametrize("name", ["trna", "score.1"])
def test_prune_option_equals_rnamotif_piped_through_rmprune(name):
    """GPUMOTIF_PRUNE=1 (gm_prune_hits on the accepted hits' records) prints
    byte for byte what `rnamotif ... | rmprune` prints.  The host side of this is also
    checked without a GPU in tests/test_host_driver_cpu.py."""
    ref = run(REF_BIN, name)
    want = subprocess.run([PRUNE_BIN], input=ref, capture_output=True, timeout=600, check=True).stdout
    got = run(GPU_BIN, name, {"GPUMOTIF_PRUNE": "1"})
    assert got == want


RMFMT = os.path.join(helpers.REF, "rmfmt")


@pytest.mark.skipif(not (have and os.path.exists(RMFMT)), reason="oracle/_ref (rmfmt, test data) or rnamotif_gpu not built")
@pytest.mark.parametrize("name", sorted(MD5))
def test_make_test_chk_goldens(name):
    """The reference's own `make test` (test/Makefile:30-245): rnamotif ... | rmfmt -l
    diffed against test/<name>.chk -- with rnamotif_gpu in rnamotif's place and the
    reference's rmfmt and .chk files as they are."""
    out = run(GPU_BIN, name)
    fmt = subprocess.run([RMFMT, "-l"], input=out, capture_output=True, timeout=600, cwd=os.path.join(DATA, "test"))
    assert fmt.returncode == 0, fmt.stderr.decode(errors="replace")[-500:]
    with open(os.path.join(DATA, "test", name + ".chk"), "rb") as fh:
        assert fmt.stdout == fh.read(), f"{name}: differs from the reference's {name}.chk"


@pytest.mark.skipif(not have, reason="oracle/_ref (test data) or rnamotif_gpu not built")
@pytest.mark.parametrize("name", sorted(MD5))
def test_fmt_option_chk_goldens(name):
    """GPUMOTIF_FMT=l: rnamotif_gpu prints what `rnamotif ... | rmfmt -l` prints (gm_rmfmt):
    the reference's 24 .chk files with no reference program in the pipe."""
    out = run(GPU_BIN, name, {"GPUMOTIF_FMT": "l"})
    with open(os.path.join(DATA, "test", name + ".chk"), "rb") as fh:
        assert out == fh.read(), f"{name}: differs from the reference's {name}.chk"
