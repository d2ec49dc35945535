"""CPU: the HOST logic of the reference-side driver (rnamotif_b200/host/rm_gpu_main.c:
batching of records, sharding of a batch into start ranges with the ordered merge,
replay of the sink's tail through the reference's own score program and printer),
byte for byte against the reference's stdout.

oracle/_ref/rnamotif_hostcheck is that driver linked against oracle/gm_mock.c -- a
stand-in for the C ABI whose candidates come from the oracle port -- instead of
libgpumotif.so (test infrastructure; the product has no CPU search path).  The same
command lines run on the device in tests/test_gpu_stdout.py."""
import hashlib
import json
import os
import subprocess

import pytest

import helpers

BIN = os.path.join(helpers.REF, "rnamotif_hostcheck")
DATA = os.path.join(helpers.REF, "data")
MD5 = json.load(open(os.path.join(helpers.GOLDEN, "make_test_md5.json")))
have = os.path.exists(BIN) and os.path.exists(os.path.join(DATA, "test", "gbrna.111.0.fastn"))


def run(name, extra_env=None):
    flags = ["-sh", "-context", "-Dctx_maxlen=5"] if name.endswith(".strict") else []
    env = dict(os.environ, EFNDATA=os.path.join(DATA, "efndata"))
    env.update(extra_env or {})
    r = subprocess.run([BIN, *flags, "-descr", name + ".descr", "gbrna.111.0.fastn"],
                       cwd=os.path.join(DATA, "test"), env=env, capture_output=True, timeout=600)
    assert r.returncode == 0, r.stderr.decode(errors="replace")[-2000:]
    return r.stdout


@pytest.mark.skipif(not have, reason="oracle/_ref/rnamotif_hostcheck not built")
@pytest.mark.parametrize("name", ["trna", "trna.strict", "score.1", "score.2.strict", "efn.strict",
                                  "sprintf", "bulge", "mp.ends.strict", "nanlin"])
def test_host_driver_stdout_md5(name):
    assert hashlib.md5(run(name)).hexdigest() == MD5[name]["md5"], f"{name}: stdout differs from the reference"


@pytest.mark.skipif(not have, reason="oracle/_ref/rnamotif_hostcheck not built")
@pytest.mark.parametrize("name", ["trna", "score.1"])
def test_host_driver_small_batches_and_shards(name):
    """Records split over many uploads (score state carried across batches) and every
    batch cut into three start ranges whose candidate lists are merged before the
    stateful score replay."""
    out = run(name, {"GPUMOTIF_BATCH_NT": "300000", "GPUMOTIF_DEVICES": "0,0,0"})
    assert hashlib.md5(out).hexdigest() == MD5[name]["md5"]


REF_BIN = os.path.join(helpers.REF, "rnamotif")
PRUNE_BIN = os.path.join(helpers.REF, "rmprune")


@pytest.mark.skipif(not (have and os.path.exists(REF_BIN) and os.path.exists(PRUNE_BIN)),
                    reason="oracle/_ref (rnamotif, rmprune, rnamotif_hostcheck) not built")
@pytest.mark.parametrize("name,env", [("trna", {}), ("score.1", {}), ("mp.ends", {}), ("efn", {}),
                                      ("trna", {"GPUMOTIF_BATCH_NT": "300000", "GPUMOTIF_DEVICES": "0,0"})])
def test_prune_option_equals_rnamotif_piped_through_rmprune(name, env):
    """GPUMOTIF_PRUNE=1: the driver captures the hits its score program accepts, lets
    gm_prune_hits take rmprune's decision on their records and prints the kept ones --
    byte for byte what `rnamotif ... | rmprune` prints (score.1 and mp.ends REJECT
    candidates first; only printed hits take part in the pruning)."""
    e = dict(os.environ, EFNDATA=os.path.join(DATA, "efndata"))
    raw = subprocess.run([REF_BIN, "-descr", name + ".descr", "gbrna.111.0.fastn"], cwd=os.path.join(DATA, "test"),
                         env=e, capture_output=True, timeout=600, check=True).stdout
    want = subprocess.run([PRUNE_BIN], input=raw, capture_output=True, timeout=600, check=True).stdout
    got = run(name, dict(env, GPUMOTIF_PRUNE="1"))
    assert got == want
    assert len(want) < len(raw) or name == "nanlin"


@pytest.mark.skipif(not have, reason="oracle/_ref/rnamotif_hostcheck not built")
@pytest.mark.parametrize("name", ["trna", "trna.strict", "score.1", "score.2.strict", "efn.strict", "sprintf", "bulge",
                                  "mp.ends.strict", "nanlin", "pk1", "qu+tr.strict"])
def test_fmt_option_equals_the_chk_goldens(name):
    """GPUMOTIF_FMT=l: the driver formats its own output like `rmfmt -l` (gm_rmfmt,
    src/rmfmt.c:52-376) -- compared with the reference's test/<name>.chk, i.e. with
    what `rnamotif ... | rmfmt -l` printed for the reference's authors."""
    out = run(name, {"GPUMOTIF_FMT": "l"})
    with open(os.path.join(DATA, "test", name + ".chk"), "rb") as fh:
        assert out == fh.read()


RMFMT = os.path.join(helpers.REF, "rmfmt")


@pytest.mark.skipif(not (have and os.path.exists(RMFMT)), reason="oracle/_ref (rmfmt, rnamotif_hostcheck) not built")
@pytest.mark.parametrize("opt,flag", [("1", []), ("la", ["-la"])])
def test_fmt_option_equals_rmfmt_other_modes(opt, flag):
    raw = run("efn")
    want = subprocess.run([RMFMT, *flag], input=raw, capture_output=True, timeout=600, check=True).stdout
    assert run("efn", {"GPUMOTIF_FMT": opt}) == want and len(want) > 0
