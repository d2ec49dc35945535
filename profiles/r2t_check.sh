#!/bin/bash
# profiles/r2t_check.sh -- GPU parity suite + default bench (shared-space assumptions on bitsets and lane state)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2t_pytest.txt 2>&1; tail -5 gpurun_out/r2t_pytest.txt
S=$(date +%s); timeout 900 python bench.py > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err; echo "bench exit $? in $(( $(date +%s) - S )) s"; tail -3 gpurun_out/r2t_bench.err
