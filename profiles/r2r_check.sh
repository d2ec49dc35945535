#!/bin/bash
# profiles/r2r_check.sh -- GPU parity suite, quickbench at 1 Gnt (pack kernels on their own stream), ncu capture of the headline command
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2r_pytest.txt 2>&1; tail -5 gpurun_out/r2r_pytest.txt
bash profiles/quickbench.sh r2r 1024 trna ire score.1 pk1 qu+tr 2>&1 | tee gpurun_out/r2r_quick.txt
bash profiles/capture_r2b.sh r2c > gpurun_out/r2c_capture.log 2>&1; tail -3 gpurun_out/r2c_capture.log
