"""FASTA ("fastn") record reader with the semantics of the reference's
FN_fgetseq (src/dbutil.c:42-128): a record starts at '>', the id is the first
blank-delimited token, the rest of the line is the definition, and the sequence
is every alphabetic character up to the next '>' ANYWHERE (not only at line
start), lower-cased, with u -> t.
"""
from __future__ import annotations

import numpy as np

_ALPHA = np.zeros(256, dtype=bool)
for _c in range(256):
    _ALPHA[_c] = chr(_c).isalpha() if _c < 128 else False
_LOWER = np.arange(256, dtype=np.uint8)
for _c in range(ord("A"), ord("Z") + 1):
    _LOWER[_c] = _c + 32
_LOWER[ord("U")] = ord("t")
_LOWER[ord("u")] = ord("t")


def parse_fastn(data: bytes, maxslen: int = 30000001):
    """Return (ids, defs, seq_bytes, rec_off): rec_off has n+1 int64 entries
    into the concatenated lower-case sequence buffer `seq_bytes` (uint8)."""
    ids, defs, chunks, offs = [], [], [], [0]
    pos, n, total = 0, len(data), 0
    while pos < n:
        if data[pos:pos + 1] != b">":
            break  # "fastn file does not begin with '>'" ends the input
        eol = data.find(b"\n", pos)
        if eol < 0:
            eol = n
        hdr = data[pos + 1:eol].lstrip(b" \t\r\f\v")
        parts = hdr.split(None, 1)
        if not parts:
            break  # unnamed entry
        sid = parts[0]
        sdef = parts[1].lstrip() if len(parts) > 1 else b""
        nxt = data.find(b">", eol)
        if nxt < 0:
            nxt = n
        body = np.frombuffer(data, dtype=np.uint8, count=max(nxt - eol, 0), offset=min(eol, n))
        seq = _LOWER[body[_ALPHA[body]]]
        if seq.size > maxslen - 1:
            seq = seq[:maxslen - 1]
        ids.append(sid.decode("latin-1"))
        defs.append(sdef.decode("latin-1"))
        chunks.append(seq)
        total += int(seq.size)
        offs.append(total)
        pos = nxt
    seq_all = np.concatenate(chunks) if chunks else np.zeros(0, dtype=np.uint8)
    return ids, defs, np.ascontiguousarray(seq_all, dtype=np.uint8), np.asarray(offs, dtype=np.int64)


def read_fastn(path: str, maxslen: int = 30000001):
    with open(path, "rb") as fh:
        return parse_fastn(fh.read(), maxslen)
