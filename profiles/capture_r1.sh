#!/bin/bash
# profiles/capture_r1.sh -- the ncu evidence of the round, taken on the default bench command
# (trna, 1024 Mnt, one B200).  Every ncu run follows a plain run of the same command that exited 0.
set -x
CMD="python bench.py --steps 2 --warmup 3 --no-cpu"
timeout 120 $CMD > gpurun_out/cap_plain.log 2>&1 || exit 1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1b.csv $CMD > gpurun_out/cap_launches.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:gm_search_kernel -s 3 -c 1 -f -o gpurun_out/prof_r1b_sieve $CMD > gpurun_out/cap_sieve.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:gm_dfs_kernel -s 3 -c 1 -f -o gpurun_out/prof_r1b_dfs $CMD > gpurun_out/cap_dfs.log 2>&1
ls -la gpurun_out/*r1b*
