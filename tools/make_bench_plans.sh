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
