#!/bin/bash
# profiles/r2v_check.sh -- final check of the round: GPU parity suite, default bench (every command under its own timeout)
mkdir -p gpurun_out
timeout 420 python -m pytest tests -m gpu -x -q > gpurun_out/r2v_pytest.txt 2>&1; tail -4 gpurun_out/r2v_pytest.txt
S=$(date +%s); timeout 300 python bench.py > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err; echo "bench exit $? in $(( $(date +%s) - S )) s"; tail -3 gpurun_out/r2v_bench.err
