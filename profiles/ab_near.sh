#!/bin/bash
# A/B: pair-bitset words of near span offsets kept in registers (lib_near) vs base; qu+tr on the worklist path; 16 upload chunks
bash profiles/ab_libs.sh 1024 "trna pk_j1+2 descr.trna.general" lib_base.so lib_near.so
echo "== qu+tr split path"; GPUMOTIF_PATH=split bash profiles/quickbench.sh qs 1024 qu+tr
echo "== 16 upload chunks"; GPUMOTIF_CHUNKS=16 bash profiles/quickbench.sh c16 1024 trna score.1
