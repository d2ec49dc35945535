#!/bin/bash
# profiles/r2q_check.sh -- GPU parity suite, quickbench at 1 Gnt (new tile loader)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2q_pytest.txt 2>&1; tail -5 gpurun_out/r2q_pytest.txt
bash profiles/quickbench.sh r2q 1024 trna ire score.1 pk1 pk_j1+2 qu+tr descr.trna.general 2>&1 | tee gpurun_out/r2q_quick.txt
GPUMOTIF_PATH=fused bash profiles/quickbench.sh r2qf 1024 trna score.1 2>&1 | tee -a gpurun_out/r2q_quick.txt
