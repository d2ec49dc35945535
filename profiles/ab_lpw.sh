#!/bin/bash
# profiles/ab_lpw.sh -- enumeration kernel with 16 instead of 32 state-owning lanes per warp (GPUMOTIF_DFS_LPW=16): parity, then 1 Gnt quickbench
mkdir -p gpurun_out
GPUMOTIF_DFS_LPW=16 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or both_paths or chunk_streamed" > gpurun_out/lpw_pytest.txt 2>&1; tail -3 gpurun_out/lpw_pytest.txt
for l in 32 16; do echo "== GPUMOTIF_DFS_LPW=$l"; GPUMOTIF_DFS_LPW=$l bash profiles/quickbench.sh lpw$l 1024 pk1 pk_j1+2 descr.trna.general trna qu+tr; done
