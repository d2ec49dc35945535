"""-m gpu: the CUDA search (through the C ABI of libgpumotif.so) against the
oracle port on the same seeded inputs, and against the committed golden
candidate streams of the reference binary.  Bar: bit-exact (integer work)."""
import numpy as np
import pytest

from oracle import oracle_port
from rnamotif_b200 import gpumotif, synth
import helpers

pytestmark = pytest.mark.gpu

NAMES = helpers.golden_names()


@pytest.fixture(scope="module")
def db():
    return synth.golden_db()


@pytest.mark.parametrize("name", NAMES)
def test_gpu_matches_reference_golden(name, db):
    ids, seq, off = db
    plan = helpers.load_plan(name)
    why = gpumotif.plan_check(plan)
    if why is not None:
        pytest.skip("plan not supported on the device: " + why)
    ms = gpumotif.MotifSearch(plan)
    hits = ms.find_motif(seq, off)
    st = ms.stats()
    ms.close()
    assert st.n_launches >= 1
    head, els = helpers.hits_to_rows(hits)
    ghead, gels, gctx = helpers.load_cands(name)
    assert len(head) == len(ghead), f"{name}: GPU {len(head)} vs reference {len(ghead)} candidates"
    if len(head):
        assert (head == ghead[:, :3]).all(), f"{name}: (rec, comp, szero) differ"
        assert (els == gels).all(), f"{name}: element assignments differ"
        if gctx.shape[1] == 4:
            assert (helpers.ctx_rows(hits) == gctx).all(), f"{name}: context differs"


@pytest.mark.parametrize("name", ["trna", "pk1", "qu+tr", "pk_j1+2", "score.1", "nanlin", "getbest.strict"])
@pytest.mark.parametrize("seed", [1, 2])
def test_gpu_matches_oracle_random(name, seed):
    plan = helpers.load_plan(name)
    rng = np.random.default_rng(1000 + seed)
    lengths = list(rng.integers(0, 3000, size=40)) + [70000]
    ids, seq, off = synth.random_records(seed, lengths, planted=True, iupac_rate=0.002)
    both = bool(gpumotif.plan_field(plan, 8))
    ref, _ = oracle_port.scan_db(plan, seq, off, both)
    ms = gpumotif.MotifSearch(plan)
    hits = ms.find_motif(seq, off)
    ms.close()
    helpers.assert_same_hits(hits, ref, name)


def test_tile_and_range_invariance(db):
    """Halo chunking: any tile size and any split of the start range gives the
    same candidates as one pass (SURVEY section 5 'long sequences')."""
    ids, seq, off = db
    plan = helpers.load_plan("trna")
    ms = gpumotif.MotifSearch(plan)
    ms.upload(seq, off)
    base = ms.scan()
    for tile in (32, 100, 777, 4096):
        ms.set_tile(tile)
        helpers.assert_same_hits(ms.scan(), base, f"tile {tile}")
    ms.set_tile(0)
    total = ms.total_nt
    cuts = [0, 1, 63, 5000, 77777, total]
    parts = [ms.scan(a, b) for a, b in zip(cuts[:-1], cuts[1:])]
    merged = np.concatenate(parts)
    order = np.lexsort((merged["seq"], merged["szero"], merged["comp"], merged["rec"]))
    helpers.assert_same_hits(merged[order], base, "range split")
    ms.close()


def test_hit_buffer_overflow_is_recovered(db):
    ids, seq, off = db
    plan = helpers.load_plan("getbest")
    ms = gpumotif.MotifSearch(plan)
    ms.upload(seq, off)
    base = ms.scan()
    assert len(base) > 64
    ms.set_hit_capacity(16)
    again = ms.scan()
    assert ms.stats().n_retries >= 1
    helpers.assert_same_hits(again, base, "after overflow")
    ms.close()


def test_empty_and_tiny_inputs():
    plan = helpers.load_plan("trna")
    ms = gpumotif.MotifSearch(plan)
    assert len(ms.find_motif(np.zeros(0, np.uint8), np.array([0], np.int64))) == 0
    assert len(ms.find_motif(np.frombuffer(b"acgu", np.uint8), np.array([0, 0, 4, 4], np.int64))) == 0
    ms.close()


@pytest.mark.parametrize("path", ["fused", "split"])
@pytest.mark.parametrize("name", ["trna", "descr.trna.general", "score.1", "pk1", "qu+tr", "ire", "mp.ends.strict"])
def test_both_paths_match_oracle(name, path, monkeypatch):
    """The fused kernel and the worklist pair (filter kernel -> enumeration kernel)
    must both give the oracle's candidate stream, whichever the library would pick
    for the plan (GPUMOTIF_PATH is read when the context is configured)."""
    monkeypatch.setenv("GPUMOTIF_PATH", path)
    plan = helpers.load_plan(name)
    rng = np.random.default_rng(77)
    lengths = list(rng.integers(0, 2500, size=30)) + [120000, 1, 64]
    ids, seq, off = synth.random_records(11, lengths, planted=True, iupac_rate=0.002)
    both = bool(gpumotif.plan_field(plan, 8))
    ref, _ = oracle_port.scan_db(plan, seq, off, both)
    ms = gpumotif.MotifSearch(plan)
    hits = ms.find_motif(seq, off)
    ms.close()
    helpers.assert_same_hits(hits, ref, f"{name} ({path})")


def test_worklist_overflow_is_recovered(monkeypatch):
    """Segments of the worklist path grow with the measured survivor rate; one that
    overflows the worklist after all is detected and the scan repeated with segments
    that cannot (gm_scan_finish).  Forced here with a tiny worklist."""
    monkeypatch.setenv("GPUMOTIF_PATH", "split")
    monkeypatch.setenv("GPUMOTIF_WL_CAP", "2048")
    monkeypatch.setenv("GPUMOTIF_SEG_NT", "4000000")
    plan = helpers.load_plan("score.1")
    ids, seq, off = synth.random_records(5, [150000, 3000, 90000], planted=True)
    both = bool(gpumotif.plan_field(plan, 8))
    ref, _ = oracle_port.scan_db(plan, seq, off, both)
    ms = gpumotif.MotifSearch(plan)
    hits = ms.find_motif(seq, off)
    retries = ms.stats().n_retries
    again = ms.find_motif(seq, off)  # second scan: segment size settled
    ms.close()
    assert retries >= 1, "the tiny worklist should have overflowed on the first scan"
    helpers.assert_same_hits(hits, ref, "after worklist overflow")
    helpers.assert_same_hits(again, ref, "scan after the overflow")


@pytest.mark.parametrize("knob", ["GPUMOTIF_NO_SIEVE", "GPUMOTIF_NO_DEEP", "GPUMOTIF_NO_DEEP2", "GPUMOTIF_NO_TAIL",
                                  "GPUMOTIF_NO_LITERAL", "GPUMOTIF_HOST_SORT", "GPUMOTIF_NO_LOOK", "GPUMOTIF_NO_PROBE",
                                  "GPUMOTIF_NO_CHAIN"])
@pytest.mark.parametrize("name", ["trna", "ire", "pk1", "pk_j1+2", "qu+tr"])
def test_filters_are_output_neutral(name, knob, monkeypatch):
    """Every level-0 filter and look-ahead only prunes what cannot reach the hit
    sink, and the device-side ordering equals the host's: with any of them switched
    off the candidate stream is still the oracle's."""
    monkeypatch.setenv(knob, "1")
    plan = helpers.load_plan(name)
    ids, seq, off = synth.random_records(23, [40000, 700, 0, 9000, 65000], planted=True, iupac_rate=0.003)
    both = bool(gpumotif.plan_field(plan, 8))
    ref, _ = oracle_port.scan_db(plan, seq, off, both)
    ms = gpumotif.MotifSearch(plan)
    hits = ms.find_motif(seq, off)
    ms.close()
    helpers.assert_same_hits(hits, ref, f"{name} with {knob}")


@pytest.mark.parametrize("name", ["trna", "score.1", "pk_j1+2", "ire"])
def test_start_count_matches_oracle(name):
    """gm_scan_stats_t::n_starts = (start, strand) pairs RM_find_motif would search
    (szero in [0, slen - rm_dminlen], src/find_motif.c:184-205): counted by the kernel
    where it visits starts one by one, by the host where the sieve does not."""
    plan = helpers.load_plan(name)
    ids, seq, off = synth.random_records(31, [5000, 0, 12, 62, 63, 64, 100, 30000, 7], planted=False)
    both = bool(gpumotif.plan_field(plan, 8))
    _, ost = oracle_port.scan_db(plan, seq, off, both)
    ms = gpumotif.MotifSearch(plan)
    ms.find_motif(seq, off)
    n = ms.stats().n_starts
    ms.close()
    assert n == ost.n_starts, f"{name}: {n} starts vs oracle {ost.n_starts}"


@pytest.mark.parametrize("seg", ["default", "small"])
@pytest.mark.parametrize("name", ["trna", "descr.trna.general", "pk1", "pk_j1+2", "qu+tr", "score.1"])
def test_chunk_streamed_scan_matches_oracle(name, seg, monkeypatch):
    """The first scan after an upload runs chunk by chunk behind the copy (launch():
    stream_in), on the worklist path with the enumeration deferred to one launch at
    the end (defer_dfs) unless the range spans several segments; later scans of the
    same upload take the plain path with a segment size grown from the measured
    survivor rate.  At bench sizes these paths carry the headline number; here they
    are forced on a small input (GPUMOTIF_CHUNK_NT) and checked against the oracle."""
    monkeypatch.setenv("GPUMOTIF_CHUNK_NT", "16384")
    if seg == "small":
        monkeypatch.setenv("GPUMOTIF_SEG_NT", "50000")
    plan = helpers.load_plan(name)
    rng = np.random.default_rng(5)
    lengths = list(rng.integers(0, 6000, size=40)) + [150000, 3, 70000]
    ids, seq, off = synth.random_records(17, lengths, planted=True, iupac_rate=0.002)
    both = bool(gpumotif.plan_field(plan, 8))
    ref, _ = oracle_port.scan_db(plan, seq, off, both)
    ms = gpumotif.MotifSearch(plan)
    first = ms.find_motif(seq, off)          # fresh upload: chunk-streamed
    n_first = ms.stats().n_launches
    second = ms.scan()                        # same upload again: plain path, grown segments
    third = ms.find_motif(seq, off)           # and a fresh upload with the settled segment size
    ms.close()
    assert n_first >= 2, "the upload should have been cut into several chunks"
    helpers.assert_same_hits(first, ref, f"{name}: chunk-streamed scan")
    helpers.assert_same_hits(second, ref, f"{name}: rescan")
    helpers.assert_same_hits(third, ref, f"{name}: second upload")


def test_two_contexts_with_different_plans_interleaved():
    """Contexts share nothing (each has its own device copy of the plan): scans of two
    different descriptors launched back to back on one device, finished in the other
    order, both give their oracle's stream."""
    ids, seq, off = synth.random_records(41, [60000, 500, 90000, 12], planted=True, iupac_rate=0.002)
    names = ["trna", "pk1", "qu+tr"]
    plans = [helpers.load_plan(n) for n in names]
    refs = [oracle_port.scan_db(p, seq, off, bool(gpumotif.plan_field(p, 8)))[0] for p in plans]
    ctxs = [gpumotif.MotifSearch(p) for p in plans]
    for _ in range(2):
        for c in ctxs:
            c.upload(seq, off)
        for c in ctxs:
            c.scan_launch()
        for c, r, n in reversed(list(zip(ctxs, refs, names))):
            c.scan_finish()
            helpers.assert_same_hits(c.hits(), r, f"{n} interleaved")
    for c in ctxs:
        c.close()


def test_setters_refuse_while_a_scan_is_in_flight():
    ids, seq, off = synth.random_records(3, [30000], planted=True)
    ms = gpumotif.MotifSearch(helpers.load_plan("trna"))
    ms.upload(seq, off)
    ms.scan_launch()
    with pytest.raises(gpumotif.GpuMotifError):
        ms.set_tile(256)
    with pytest.raises(gpumotif.GpuMotifError):
        ms.set_hit_capacity(1 << 12)
    ms.scan_finish()
    ms.set_tile(256)
    ms.close()


def test_hit_dense_input_and_long_windows():
    """Descriptors of the reference's corpus whose candidate volume or window kept
    them out of the committed goldens (tests/golden/manifest.json), on inputs small
    enough for the oracle: mpr / phlx.pfrac (mispair-tolerant hairpins: more than one
    candidate per nucleotide), hlx.gf.iu (1000-nt loops: 50 candidates per
    nucleotide), pk.gf.iu (pseudoknot with three 1000-nt single strands: 10^5
    candidates from a few hundred nucleotides), eloop (the 6000-nt default window)."""
    for name, n in (("descr.mpr", 6000), ("descr.hlx.gf.iu", 9000), ("descr.phlx.pfrac", 9000),
                    ("descr.pk.gf.iu", 400), ("descr.eloop", 8000)):
        plan = helpers.load_extra_plan(name)
        if plan is None:
            pytest.skip("plan of %s not committed" % name)
        ids, seq, off = synth.random_records(9, [n, 300, n // 2], planted=True, iupac_rate=0.001)
        both = bool(gpumotif.plan_field(plan, 8))
        ref, _ = oracle_port.scan_db(plan, seq, off, both, cap_hits=1 << 22)
        ms = gpumotif.MotifSearch(plan)
        hits = ms.find_motif(seq, off)
        ms.close()
        helpers.assert_same_hits(hits, ref, name)
        assert len(ref) > 0, name


@pytest.mark.parametrize("name,nt", [("ire", 30_000_000), ("score.1", 5_490_000)])
def test_one_very_long_record(name, nt):
    """A single record of chromosome size (the reference's MAXSLEN is 30 000 000,
    src/rnamot.h): tiles and worklist segments all fall inside one record, starts on
    the complementary strand count from its far end."""
    plan = helpers.load_plan(name)
    ids, seq, off = synth.random_records(101, [nt, 70], planted=False)
    both = bool(gpumotif.plan_field(plan, 8))
    ref, _ = oracle_port.scan_db(plan, seq, off, both)
    ms = gpumotif.MotifSearch(plan)
    hits = ms.find_motif(seq, off)
    ms.close()
    helpers.assert_same_hits(hits, ref, f"{name}: one {nt}-nt record")
    assert len(ref) > 0
