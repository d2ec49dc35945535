#!/bin/bash
# profiles/r2w_check.sh -- final check of the round (after the lite lane-window build): GPU parity suite, default bench (every command under its own timeout)
mkdir -p gpurun_out
timeout 420 python -m pytest tests -m gpu -x -q > gpurun_out/r2w_pytest.txt 2>&1; tail -4 gpurun_out/r2w_pytest.txt
S=$(date +%s); timeout 300 python bench.py > gpurun_out/r2w_bench.json 2> gpurun_out/r2w_bench.err; echo "bench exit $? in $(( $(date +%s) - S )) s"; tail -3 gpurun_out/r2w_bench.err
