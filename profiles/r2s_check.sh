#!/bin/bash
# profiles/r2s_check.sh -- GPU parity suite + 1 Gnt quickbench (shared-space assumptions: generic LD.E -> LDS)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2s_pytest.txt 2>&1; tail -5 gpurun_out/r2s_pytest.txt
bash profiles/quickbench.sh r2s 1024 trna ire score.1 pk1 pk_j1+2 qu+tr descr.trna.general 2>&1 | tee gpurun_out/r2s_quick.txt
