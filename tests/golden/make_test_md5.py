#!/usr/bin/env python3
"""Record md5(raw rnamotif stdout) of the reference build for the 24 `make test`
command lines (test/Makefile:30-245) into tests/golden/make_test_md5.json.
oracle/check_goldens.sh verifies the same runs against the reference's .chk
files first.  Usage: oracle/check_goldens.sh | python tests/golden/make_test_md5.py"""
import json
import os
import sys

out = {}
for ln in sys.stdin:
    name, verdict, md5, hits = ln.split()
    assert verdict == "PASS", ln
    out[name] = {"md5": md5, "hits": int(hits)}
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "make_test_md5.json"), "w"), indent=1)
