"""The score section's pre-screen (include/gpumotif_score.h,
rnamotif_b200/csrc/gm_score.h): candidates the MAIN score program REJECTs outright are
dropped at the device's hit sink and never replayed on the host.

CPU: the same interpreter the device runs (gm_score_prescreen, host build of
gm_score.h) over the committed candidate streams of the instrumented reference --
every candidate carries RM_score's verdict (tests/golden/cands: action 0 = REJECT):
the pre-screen never drops a candidate the reference accepts, and for programs
that only filter (score.1, score.2, mp.ends, trna.general) it drops exactly the
reference's rejects.  Programs with state between candidates are refused.
GPU: the device's kept set equals the reference's accepted set."""
import ctypes as C

import numpy as np
import pytest

import helpers
from oracle import oracle_port
from rnamotif_b200 import gpumotif, synth

PURE = ["score.1", "score.2", "mp.ends", "descr.trna.general", "score.1.strict", "mp.ends.strict"]
MIXED = ["ire", "efn", "sprintf", "bulge", "descr.score.0", "descr.score.3"]


def score_of(name):
    return helpers.load_score(name[:-7] if name.endswith(".strict") else name)


def strand(seq, off, rec, comp):
    s = bytes(seq[off[rec]:off[rec + 1]]).lower().replace(b"u", b"t")
    if comp:
        s = s[::-1].translate(bytes.maketrans(b"acgt", b"tgca") if False else _RC)
    return s


_RC = bytes((ord("n") if chr(c) not in "acgt" else ord("tgca"["acgt".index(chr(c))])) for c in range(256))


@pytest.fixture(scope="module")
def db():
    return synth.golden_db()


@pytest.mark.parametrize("name", PURE + MIXED)
def test_host_prescreen_against_reference_verdicts(name, db):
    ids, seq, off = db
    plan, score = helpers.load_plan(name), score_of(name)
    assert score is not None and helpers.score_present(score), name
    both = bool(gpumotif.plan_field(plan, 8))
    hits, _ = oracle_port.scan_db(plan, seq, off, both)
    ghead, gels, _ = helpers.load_cands(name)
    assert len(hits) == len(ghead)
    L = gpumotif.lib()
    rej = np.zeros(len(hits), dtype=bool)
    cache = {}
    for i in range(len(hits)):
        key = (int(hits["rec"][i]), int(hits["comp"][i]))
        if key not in cache:
            cache[key] = strand(seq, off, *key)
        sb = cache[key]
        rej[i] = bool(L.gm_score_prescreen(plan, score, hits[i:i + 1].ctypes.data, sb, len(sb)))
    accepted = ghead[:, 3] != 0
    assert not (rej & accepted).any(), f"{name}: the pre-screen drops a candidate the reference accepts"
    if name in PURE:
        assert (rej == ~accepted).all(), f"{name}: {int(rej.sum())} dropped, the reference rejects {int((~accepted).sum())}"
        assert rej.any() or name.endswith(".strict")


def test_stateful_programs_are_refused():
    sc = helpers.load_score("getbest")   # HOLD / RELEASE and an END section (test/getbest.descr:18-55)
    assert sc is not None and not helpers.score_present(sc)
    sc4 = helpers.load_score("descr.score.4")
    assert sc4 is not None and not helpers.score_present(sc4)


@pytest.mark.gpu
@pytest.mark.parametrize("name", PURE + ["ire", "efn"])
def test_device_prescreen_keeps_exactly_what_the_reference_accepts(name, db):
    ids, seq, off = db
    plan, score = helpers.load_plan(name), score_of(name)
    ghead, gels, _ = helpers.load_cands(name)
    ms = gpumotif.MotifSearch(plan)
    ms.set_score(score)
    hits = ms.find_motif(seq, off)
    st = ms.stats()
    ms.set_score(None)
    allhits = ms.find_motif(seq, off)
    ms.close()
    assert len(allhits) == len(ghead) and st.n_score_rejected + len(hits) == len(ghead)
    head, els = helpers.hits_to_rows(hits)
    accepted = ghead[:, 3] != 0
    if name in PURE:
        assert len(hits) == int(accepted.sum()), f"{name}: kept {len(hits)}, the reference accepts {int(accepted.sum())}"
        assert (head == ghead[accepted][:, :3]).all() and (els == gels[accepted]).all()
    else:
        # everything the reference accepts is still there, in order
        keep = {tuple(r) for r in np.concatenate([head, els], axis=1).tolist()}
        want = np.concatenate([ghead[accepted][:, :3], gels[accepted]], axis=1).tolist()
        assert all(tuple(r) in keep for r in want)
