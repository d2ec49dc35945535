#!/bin/bash
# profiles/refill_sweep.sh -- lanes a warp of the enumeration kernel waits for before it builds new windows (GPUMOTIF_REFILL)
for r in 2 4 8 16 24; do
  echo "== GPUMOTIF_REFILL=$r"
  GPUMOTIF_REFILL=$r bash profiles/quickbench.sh refill$r 512 pk1 pk_j1+2 descr.trna.general trna
done
