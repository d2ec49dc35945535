"""The FASTA reader (FN_fgetseq, src/dbutil.c:42-128).

CPU: the host mirror rnamotif_b200/fastn.py against the REFERENCE's own reader
(oracle/_ref/libreffastn.so = src/dbutil.c + oracle/fgetseq_hook.c, where built).
GPU: gm_db_upload_fastn (the reader on the device) against both, then the search
and the hit windows through it against the upload_chars path.
"""
import ctypes as C
import os

import numpy as np
import pytest

import helpers
from rnamotif_b200 import fastn, synth

REFLIB = os.path.join(helpers.REF, "libreffastn.so")


def ref_parse(text: bytes, maxslen: int = 30000001):
    """The reference's FN_fgetseq over `text`: (ids, defs, seq, rec_off)."""
    L = C.CDLL(REFLIB)
    L.gmo_ref_fastn.restype = C.c_int
    L.gmo_ref_fastn.argtypes = [C.c_char_p, C.c_long, C.c_int, C.c_void_p, C.c_long, C.c_void_p, C.c_void_p,
                                C.c_long, C.c_int]
    n = len(text)
    max_rec = text.count(b">") + 1
    seq = np.zeros(n + 1, dtype=np.uint8)
    off = np.zeros(max_rec + 1, dtype=np.int64)
    ids = C.create_string_buffer(2 * n + 64 * max_rec + 64)
    assert C.sizeof(C.c_long) == 8
    r = L.gmo_ref_fastn(text, n, maxslen, seq.ctypes.data, n + 1, off.ctypes.data, ids, len(ids), max_rec)
    assert r >= 0
    rows = [ln.split("\t", 1) for ln in ids.value.decode("latin-1").split("\n")[:-1]]
    return [a for a, _ in rows], [b for _, b in rows], seq[:off[r]].copy(), off[:r + 1].copy()


def golden_text() -> bytes:
    ids, seq, off = synth.golden_db()
    out = []
    for i, sid in enumerate(ids):
        out.append(b">" + sid.encode() + b" synthetic record %d\n" % i)
        s = seq[off[i]:off[i + 1]].tobytes()
        for k in range(0, len(s), 70):
            out.append(s[k:k + 70] + b"\n")
    return b"".join(out)


# (name, text): what the reader has to get right
CASES = [
    ("plain", b">a first\nACGU\nacgt\n>b\nGGCC\n"),
    ("no_final_newline", b">a\nacgt\n>b x\nggcc"),
    ("gt_inside_header", b">a has > inside > the header\nacgt\n>b\ntt\n"),
    ("gt_mid_line", b">a\nacgt>b second starts mid line\ncc\ngg>c\n\n>d\nn\n"),
    ("non_alpha_dropped", b">a\nac gt\t12 3*-.\r\nNNrykm\n>b\n  \n>c\n1234\n"),
    ("iupac_and_other_letters", b">a\nacgturykmswbdhvnxzjqACGTURYKMSWBDHVNXZJQ\n"),
    ("blank_lines", b">a\n\n\nac\n\n\ngt\n\n>b\n\n"),
    ("header_only_at_eof", b">a\nacgt\n>b only a header"),
    ("header_then_eof_newline", b">a\nacgt\n>b\n"),
    ("empty_records", b">a\n>b\n>c\nacgt\n>d\n"),
    ("crlf", b">a def\r\nacgt\r\nacgt\r\n>b\r\ngg\r\n"),
    ("spaces_before_id", b">   a   the def\nacgt\n>\tb\nacgt\n"),
    ("one_record_no_newline", b">a"),
    ("high_bytes", b">a\nac\xe9\xffgt\x80\n"),
    ("long_line", b">a\n" + b"acgu" * 5000 + b"\n>b\n" + b"g" * 33 + b"\n"),
    ("many_short", b"".join(b">r%d\n%s\n" % (i, b"acgtn"[: i % 6]) for i in range(3000))),
]


def seg_boundary_text() -> bytes:
    """Events placed on both sides of the 16 KB segment and 512 B / 16 B lane
    boundaries of the device reader."""
    rng = np.random.default_rng(7)
    parts = []
    pos = 0
    for target in [16384 - 1, 16384, 16384 + 1, 2 * 16384 - 16, 3 * 16384 - 512, 3 * 16384 + 15, 5 * 16384]:
        hdr = b">r%d a header that runs across the boundary > with a gt in it\n" % target
        # place the '>' so that the header straddles `target`
        fill = target - pos - 10
        if fill > 0:
            body = bytes(rng.choice(np.frombuffer(b"acgtACGTnu\n", dtype=np.uint8), fill))
            parts.append(body)
            pos += len(body)
        parts.append(hdr)
        pos += len(hdr)
    parts.append(b"acgt" * 10000)  # a long run without any event: whole segments with none
    return b">first\n" + b"".join(parts)


CASES.append(("segment_boundaries", seg_boundary_text()))


def mirror(text: bytes):
    ids, defs, seq, off = fastn.parse_fastn(text)
    return ids, defs, seq, off


needs_ref = pytest.mark.skipif(not os.path.exists(REFLIB), reason="oracle/_ref not built")


@needs_ref
@pytest.mark.parametrize("name,text", CASES, ids=[c[0] for c in CASES])
def test_mirror_matches_reference_reader(name, text):
    ri, rd, rs, ro = ref_parse(text)
    mi, md, ms, mo = mirror(text)
    assert mo.tolist() == ro.tolist()
    assert ms.tobytes() == rs.tobytes()
    assert mi == ri
    assert md == rd


@needs_ref
def test_mirror_matches_reference_reader_golden_db():
    text = golden_text()
    ri, rd, rs, ro = ref_parse(text)
    mi, md, ms, mo = mirror(text)
    assert mo.tolist() == ro.tolist() and ms.tobytes() == rs.tobytes() and mi == ri and md == rd
    ids, seq, off = synth.golden_db()
    assert ro.tolist() == off.tolist()


def _hdr_expected(text: bytes):
    """Offsets of the '>' that start records, by the two-state rule."""
    out, in_hdr = [], False
    for i, b in enumerate(text):
        if b == 0x3E and not in_hdr:
            out.append(i)
            in_hdr = True
        elif b == 0x0A:
            in_hdr = False
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("name,text", CASES, ids=[c[0] for c in CASES])
def test_device_reader_matches_reference_reader(name, text):
    from rnamotif_b200 import gpumotif
    ms = gpumotif.MotifSearch(helpers.load_plan("trna"))
    ms.upload_fastn(text)
    rec, hdr = ms.records()
    if os.path.exists(REFLIB):
        _, _, seq, off = ref_parse(text)
    else:
        _, _, seq, off = mirror(text)
    # the reference stops a file at an unnamed entry (src/dbutil.c:62-66); the
    # device returns every record and leaves that decision to the caller, so
    # compare on texts where it does not occur -- all of CASES
    assert rec.tolist() == off.tolist()
    assert hdr.tolist() == _hdr_expected(text) + [len(text)]
    assert ms.get_chars().tobytes() == seq.tobytes()
    ms.close()


@pytest.mark.gpu
def test_device_reader_empty_and_errors():
    from rnamotif_b200 import gpumotif
    ms = gpumotif.MotifSearch(helpers.load_plan("trna"))
    ms.upload_fastn(b"")
    rec, hdr = ms.records()
    assert rec.tolist() == [0] and hdr.tolist() == [0]
    assert len(ms.scan()) == 0
    with pytest.raises(gpumotif.GpuMotifError, match="does not begin with '>'"):
        ms.upload_fastn(b"acgt\n>a\nacgt\n")
    ms.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["trna", "score.1.strict", "pk1", "qu+tr", "extra.ctx_seq.strict"])
def test_search_through_device_reader(name):
    """Same candidates whether the records arrive as characters or as FASTA text;
    the hit windows hold what fm_sbuf would."""
    from rnamotif_b200 import gpumotif
    ids, seq, off = synth.golden_db()
    text = golden_text()
    plan = helpers.load_plan(name)
    a = gpumotif.MotifSearch(plan)
    want = a.find_motif(seq, off)
    a.close()
    b = gpumotif.MotifSearch(plan)
    b.upload_fastn(text)
    got = b.scan()
    helpers.assert_same_hits(want, got, name)
    assert helpers.ctx_rows(want).tolist() == helpers.ctx_rows(got).tolist()
    lead, trail = 7, 9
    win = b.hit_windows(lead, trail)
    b.close()
    assert win.shape[0] == len(got)
    # expected: the strand as the reference holds it
    lower = fastn._LOWER
    comp = np.full(256, ord("n"), dtype=np.uint8)
    for x, y in zip(b"acgt", b"tgca"):
        comp[x] = y
    rng = np.random.default_rng(3)
    pick = rng.choice(len(got), size=min(len(got), 400), replace=False) if len(got) else []
    for i in pick:
        r, c, z = int(got["rec"][i]), int(got["comp"][i]), int(got["szero"][i])
        s = lower[seq[off[r]:off[r + 1]]]
        if c:
            s = comp[s[::-1]]
        w = win[i]
        for j in range(win.shape[1]):
            p = z - lead + j
            exp = s[p] if 0 <= p < len(s) else 0
            assert w[j] == exp, (name, i, j)
