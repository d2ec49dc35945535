#!/bin/bash
# profiles/capture_r2.sh -- the ncu evidence of round 2.  Every ncu run follows a plain run of the same
# command that exited 0.  HEAD = the headline part of the default bench command (trna, 1024 Mnt, one B200);
# the per-config captures use the same harness at 256 Mnt (`--descr X --mnt 256`).
# The first scan of a context runs default segments of 16 Mnt (the segment size then follows the measured
# survivor rate), so at 1024 Mnt launch 65 of a kernel is its first bench-size launch, at 256 Mnt launch 17.
set -x
mkdir -p gpurun_out
HEAD="python bench.py --steps 2 --warmup 3 --no-cpu --configs none --no-parity --no-binary"
timeout 200 $HEAD > gpurun_out/cap2_plain.log 2>&1 || exit 1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r2.csv $HEAD > gpurun_out/cap2_launches.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gm_filter_kernel -s 65 -c 1 -f -o gpurun_out/prof_r2_sieve $HEAD > gpurun_out/cap2_sieve.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gm_dfs_kernel -s 65 -c 1 -f -o gpurun_out/prof_r2_dfs $HEAD > gpurun_out/cap2_dfs.log 2>&1
for d in pk1 pk_j1+2 trna.general; do
  CMD="python bench.py --descr $d --mnt 256 --steps 1 --warmup 2 --no-cpu --configs none --no-parity --no-binary"
  timeout 200 $CMD > gpurun_out/cap2_plain_$d.log 2>&1 || continue
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:gm_dfs_kernel -s 17 -c 1 -f -o gpurun_out/prof_r2_dfs_$d $CMD > gpurun_out/cap2_dfs_$d.log 2>&1
done
for d in qu+tr score.1; do
  CMD="python bench.py --descr $d --mnt 256 --steps 1 --warmup 2 --no-cpu --configs none --no-parity --no-binary"
  timeout 200 $CMD > gpurun_out/cap2_plain_$d.log 2>&1 || continue
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:gm_search_kernel -s 3 -c 1 -f -o gpurun_out/prof_r2_fused_$d $CMD > gpurun_out/cap2_fused_$d.log 2>&1
done
# summaries are made here (gpurun brings back at most 64 MiB): the .ncu-rep files stay on the box
for r in gpurun_out/prof_r2_*.ncu-rep; do
  b=$(basename $r .ncu-rep)
  python profiles/summarize.py full $r > gpurun_out/${b#prof_}.txt 2>&1
done
python profiles/summarize.py json gpurun_out/prof_r2_sieve.ncu-rep trna 1024 > gpurun_out/r2_sieve_kernel.json
python profiles/summarize.py launches gpurun_out/launches_r2.csv > gpurun_out/r2_launches.txt
rm -f gpurun_out/prof_r2_*.ncu-rep
ls -la gpurun_out/
