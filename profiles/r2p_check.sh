#!/bin/bash
# profiles/r2p_check.sh -- GPU parity suite, default bench (both upload modes in the e2e leg)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2p_pytest.txt 2>&1; tail -5 gpurun_out/r2p_pytest.txt
S=$(date +%s); timeout 900 python bench.py > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err; echo "bench exit $? in $(( $(date +%s) - S )) s"; tail -3 gpurun_out/r2p_bench.err
for t in 4 8 16; do GPUMOTIF_PACK_THREADS=$t python bench.py --descr ire --mnt 1024 --steps 3 --warmup 3 --no-cpu --configs none --no-parity --no-binary --upload hostpack > gpurun_out/r2p_ire_t$t.json 2>/dev/null; python - gpurun_out/r2p_ire_t$t.json $t <<'PY'
import json,sys
j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print("ire hostpack threads", sys.argv[2], "e2e", round(j["e2e"]["value"],1), j["e2e"]["phases_ms_last_step"])
PY
done
