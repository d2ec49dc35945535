#!/bin/bash
# tools/make_bench_plans.sh -- flatten the benchmark descriptors with the reference's own front end
# (rnamotif_b200/host/_build/rm_plan_dump = rnamot.c's compile steps + rm_flatten.c) into
# rnamotif_b200/plans/<name>.plan.gz, which bench.py loads.  Needs /root/reference and
# `make -C oracle ref && make -C rnamotif_b200/host`.
set -e
root=$(cd "$(dirname "$0")/.." && pwd)
dump=$root/rnamotif_b200/host/_build/rm_plan_dump
ref=${REFERENCE:-/root/reference}
out=$root/rnamotif_b200/plans
mkdir -p "$out"
one() { # name, descriptor path
  tmp=$(mktemp)
  (cd "$(dirname "$2")" && EFNDATA=$ref/efndata GM_PLAN_OUT=$tmp "$dump" -descr "$(basename "$2")" > /dev/null)
  gzip -n -9 -c "$tmp" > "$out/$1.plan.gz"
  rm -f "$tmp"
  echo "$1: $(zcat "$out/$1.plan.gz" | wc -c) bytes"
}
for d in trna ire score.1 pk1 pk_j1+2 qu+tr mp.ends efn; do one $d $ref/test/$d.descr; done
one trna.general $ref/Ecoli.trna.example/trna.general.descr
# plans of the corpus descriptors whose candidate volume / run time kept them out of the committed
# golden streams (tests/golden/manifest.json "skipped"): the GPU tests run them on small inputs
# against the oracle port (tests/test_gpu_parity.py::test_hit_dense_input_and_long_windows)
out=$root/tests/golden/plans_extra
mkdir -p "$out"
for d in eloop hlx.gf.iu hlx.gu.iu mpr phlx.pfrac pk.gf.iu pk.gu.iu; do one descr.$d $ref/descr/$d.descr; done
# the MAIN score programs as the device's pre-screen takes them (include/gpumotif_score.h)
out=$root/tests/golden/scores
mkdir -p "$out"
score() { # name, descriptor path
  tmp=$(mktemp); tmp2=$(mktemp)
  (cd "$(dirname "$2")" && EFNDATA=$ref/efndata GM_PLAN_OUT=$tmp GM_SCORE_OUT=$tmp2 "$dump" -descr "$(basename "$2")" 2>&1 | grep "pre-screen")
  gzip -n -9 -c "$tmp2" > "$out/$1.score.gz"
  rm -f "$tmp" "$tmp2"
}
for d in score.1 score.2 mp.ends ire efn getbest sprintf bulge; do score $d $ref/test/$d.descr; done
score descr.trna.general $ref/Ecoli.trna.example/trna.general.descr
for d in score.0 score.3 score.4 nanlin tmrna; do [ -f $ref/descr/$d.descr ] && score descr.$d $ref/descr/$d.descr; done
# ... and beside the benchmark plans
for d in score.1 ire efn mp.ends; do cp "$out/$d.score.gz" "$root/rnamotif_b200/plans/$d.score.gz"; done
cp "$out/descr.trna.general.score.gz" "$root/rnamotif_b200/plans/trna.general.score.gz"
