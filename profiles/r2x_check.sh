#!/bin/bash
# profiles/r2x_check.sh -- two-stage sieve from a literal / chain alone: GPU parity suite, quickbench of the plans it changes, default bench
mkdir -p gpurun_out
timeout 330 python -m pytest tests -m gpu -x -q > gpurun_out/r2x_pytest.txt 2>&1; tail -3 gpurun_out/r2x_pytest.txt
for k in 0 1; do echo "== GPUMOTIF_NO_TWO_STAGE=$k (pk1 ire descr.quad)"; if [ $k = 1 ]; then export GPUMOTIF_NO_TWO_STAGE=1; fi; timeout 120 bash profiles/quickbench.sh two$k 1024 pk1 ire descr.quad; done; unset GPUMOTIF_NO_TWO_STAGE
S=$(date +%s); timeout 200 python bench.py > gpurun_out/r2x_bench.json 2> gpurun_out/r2x_bench.err; echo "bench exit $? in $(( $(date +%s) - S )) s"
