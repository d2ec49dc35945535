#!/usr/bin/env python3
"""Regenerate tests/golden/ from the REFERENCE (needs /root/reference and the
binaries built by `make -C oracle ref` and `make -C rnamotif_b200/host`):

  plans/<name>.plan.gz   the flattened plan (include/gpumotif_plan.h) of each
                         descriptor, produced by the reference's own front end
                         (rnamotif_b200/host/rm_plan_dump)
  cands/<name>.npz       the candidate stream of the instrumented reference
                         binary (oracle/_ref/rnamotif_cand -O0) over
                         rnamotif_b200.synth.golden_db(): every assignment that
                         reaches the hit sink (src/find_motif.c:362-394), in
                         enumeration order, with RM_score's verdict.  The binary
                         runs with oracle/_ref/malloc_ff.so preloaded: its
                         -strict_helices checks read fm_window[] cells nothing has
                         written yet (uninitialised malloc memory, oracle/malloc_ff.c);
                         the shim makes them read UNDEF, so the stream depends on
                         the input only
  manifest.json          names, flags, candidate counts, md5 of raw stdout

Usage: python tests/golden/make_golden.py
"""
import gzip
import hashlib
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from rnamotif_b200 import synth  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref")
DUMP = os.path.join(ROOT, "rnamotif_b200", "host", "_build", "rm_plan_dump")
TESTS = ["nanlin", "pk1", "pk_j1+2", "qu+tr", "score.1", "score.2", "trna", "mp.ends", "efn", "sprintf",
         "bulge", "getbest", "ire", "ire.1"]
STRICT = ["-sh", "-context", "-Dctx_maxlen=5"]
MAX_CANDS = 150000


def run_one(name, descr_path, flags, fasta, out_plans, out_cands):
    d = os.path.dirname(descr_path)
    with tempfile.TemporaryDirectory() as tmp:
        env = dict(os.environ, EFNDATA=os.path.join(REF, "data", "efndata"),
                   GM_PLAN_OUT=os.path.join(tmp, "p.plan"), GM_CAND_FILE=os.path.join(tmp, "c.txt"))
        r = subprocess.run([DUMP, *flags, "-descr", os.path.basename(descr_path)], cwd=d, env=env,
                           capture_output=True)
        if r.returncode != 0:
            return {"name": name, "skipped": "plan: " + r.stderr.decode(errors="replace").strip().splitlines()[-1]}
        plan = open(env["GM_PLAN_OUT"], "rb").read()
        try:
            env["LD_PRELOAD"] = os.path.join(REF, "malloc_ff.so")
            r = subprocess.run([os.path.join(REF, "rnamotif_cand"), "-O0", *flags, "-descr",
                                os.path.basename(descr_path), fasta], cwd=d, env=env, capture_output=True,
                               timeout=300)
        except subprocess.TimeoutExpired:
            return {"name": name, "skipped": "reference did not finish in 300 s"}
        if r.returncode != 0:
            return {"name": name, "skipped": "reference exit %d" % r.returncode}
        nd = int(np.frombuffer(plan, dtype=np.int32, count=3)[2])
        rows = [ln.split() for ln in open(env["GM_CAND_FILE"])]
        if len(rows) > MAX_CANDS:
            return {"name": name, "skipped": "%d candidates (> %d)" % (len(rows), MAX_CANDS)}
        a = np.array(rows, dtype=np.int64).reshape(len(rows), -1) if rows else np.zeros((0, 5 + 4 * nd), np.int64)
        with gzip.GzipFile(os.path.join(out_plans, name + ".plan.gz"), "wb", mtime=0) as fh:
            fh.write(plan)
        np.savez_compressed(os.path.join(out_cands, name + ".npz"),
                            head=a[:, :4].astype(np.int32),            # rec comp szero action
                            els=a[:, 5:5 + 4 * nd].astype(np.int32),   # off len mpr mm per element
                            ctx=a[:, 5 + 4 * nd:].astype(np.int32))
        return {"name": name, "flags": list(flags), "n_descr": nd, "candidates": len(rows),
                "accepted": int((a[:, 3] != 0).sum()) if len(rows) else 0,
                "stdout_md5": hashlib.md5(r.stdout).hexdigest()}


def extras(fasta, out_plans, out_cands):
    """Purpose-written descriptors (tests/golden/extra_descr/, ours) for corners no
    shipped descriptor reaches; slack and strict."""
    edir = os.path.join(HERE, "extra_descr")
    out = []
    for f in sorted(os.listdir(edir)):
        if not f.endswith(".descr"):
            continue
        out.append(run_one("extra." + f[:-6], os.path.join(edir, f), [], fasta, out_plans, out_cands))
        print(out[-1], flush=True)
        out.append(run_one("extra." + f[:-6] + ".strict", os.path.join(edir, f), STRICT, fasta, out_plans, out_cands))
        print(out[-1], flush=True)
    return out


def main():
    out_plans = os.path.join(HERE, "plans")
    out_cands = os.path.join(HERE, "cands")
    os.makedirs(out_plans, exist_ok=True)
    os.makedirs(out_cands, exist_ok=True)
    ids, seq, off = synth.golden_db()
    manifest = []
    with tempfile.TemporaryDirectory() as tmp:
        fasta = os.path.join(tmp, "golden.fastn")
        synth.write_fastn(fasta, ids, seq, off)
        if "--plans-only" in sys.argv:
            # the plan format changed: re-dump every plan, keep the candidate streams
            old = json.load(open(os.path.join(HERE, "manifest.json")))
            for e in old["entries"]:
                if "skipped" in e:
                    continue
                name = e["name"]
                if name.startswith("extra."):
                    base, d = name[len("extra."):], os.path.join(HERE, "extra_descr")
                elif name.startswith("descr."):
                    base, d = name[len("descr."):], os.path.join(REF, "data", "descr")
                else:
                    base, d = name, os.path.join(REF, "data", "test")
                if name.startswith("extra.") and base.endswith(".strict"):
                    base = base[:-len(".strict")]
                env = dict(os.environ, GM_PLAN_OUT=os.path.join(tmp, "p.plan"))
                subprocess.run([DUMP, *e["flags"], "-descr", base + ".descr"], cwd=d, env=env, check=True,
                               capture_output=True)
                with gzip.GzipFile(os.path.join(out_plans, name + ".plan.gz"), "wb", mtime=0) as fh:
                    fh.write(open(env["GM_PLAN_OUT"], "rb").read())
            return
        if "--extra-only" in sys.argv:
            old = json.load(open(os.path.join(HERE, "manifest.json")))
            manifest = [e for e in old["entries"] if not e["name"].startswith("extra.")]
            manifest += extras(fasta, out_plans, out_cands)
            json.dump({"db": "rnamotif_b200.synth.golden_db()", "total_nt": int(off[-1]), "entries": manifest},
                      open(os.path.join(HERE, "manifest.json"), "w"), indent=1)
            return
        manifest += extras(fasta, out_plans, out_cands)
        tdir = os.path.join(REF, "data", "test")
        for t in TESTS:
            manifest.append(run_one(t, os.path.join(tdir, t + ".descr"), [], fasta, out_plans, out_cands))
            print(manifest[-1], flush=True)
            manifest.append(run_one(t + ".strict", os.path.join(tdir, t + ".strict.descr"), STRICT, fasta,
                                    out_plans, out_cands))
            print(manifest[-1], flush=True)
        ddir = os.path.join(REF, "data", "descr")
        for f in sorted(os.listdir(ddir)):
            if not f.endswith(".descr"):
                continue
            name = "descr." + f[:-6]
            manifest.append(run_one(name, os.path.join(ddir, f), [], fasta, out_plans, out_cands))
            print(manifest[-1], flush=True)
    json.dump({"db": "rnamotif_b200.synth.golden_db()", "total_nt": int(off[-1]), "entries": manifest},
              open(os.path.join(HERE, "manifest.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
