#!/bin/bash
# profiles/capture_r2c.sh -- ncu --set full of the enumeration kernel where it dominates (pk1, pk_j1+2, trna.general at 256 Mnt)
# and of the filter kernel for ire (no look-ahead passes), as shipped at the end of round 2.  Each under its own timeout.
mkdir -p gpurun_out
for d in pk1 pk_j1+2 trna.general; do
  CMD="python bench.py --descr $d --mnt 256 --steps 1 --warmup 2 --no-cpu --configs none --no-parity --no-binary --upload chars"
  timeout 120 $CMD > gpurun_out/cap3_plain_$d.log 2>&1 || continue
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:gm_dfs_kernel -s 17 -c 1 -f -o gpurun_out/prof_r2e_dfs_$d $CMD > gpurun_out/cap3_dfs_$d.log 2>&1
done
CMD="python bench.py --descr ire --mnt 256 --steps 1 --warmup 2 --no-cpu --configs none --no-parity --no-binary --upload chars"
timeout 120 $CMD > gpurun_out/cap3_plain_ire.log 2>&1 && timeout 200 ncu --set full --clock-control none --import-source on -k regex:gm_filter_kernel -s 17 -c 1 -f -o gpurun_out/prof_r2e_filter_ire $CMD > gpurun_out/cap3_filter_ire.log 2>&1
for r in gpurun_out/prof_r2e_*.ncu-rep; do
  b=$(basename $r .ncu-rep)
  timeout 120 python profiles/summarize.py full $r > gpurun_out/${b#prof_}.txt 2>&1
done
rm -f gpurun_out/prof_r2e_*.ncu-rep
ls gpurun_out | grep r2e
