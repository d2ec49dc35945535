#!/bin/bash
# profiles/quickbench.sh TAG MNT descr... -- resident/e2e throughput per descriptor (no CPU leg)
tag=$1; mnt=$2; shift 2
for d in "$@"; do
  python bench.py --descr "$d" --mnt $mnt --steps 3 --warmup 3 --no-cpu --configs none --no-parity --no-binary > gpurun_out/qb_${tag}_$d.json 2> gpurun_out/qb_${tag}_$d.err
  python - "$d" gpurun_out/qb_${tag}_$d.json <<'PY'
import json,sys
try:
    j=json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    print("%-22s value %8.3f  e2e %8.3f  kernel_ms %9.3f  cands %d" % (sys.argv[1], j["value"], j["e2e"]["value"], j["roofline"]["kernel_ms"], j["config"]["candidates_per_step_rank0"]))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
