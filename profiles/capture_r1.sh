#!/bin/bash
# profiles/capture_r1.sh -- the ncu evidence of the round, taken on the default bench command
# (trna, 1024 Mnt, one B200).  Every ncu run follows a plain run of the same command that exited 0.
# The first scan of a context runs 64 default segments of 16 Mnt (the segment size then follows the
# measured survivor rate), so launch 66 of the sieve kernel is the first bench-size one (1024 Mnt).
set -x
CMD="python bench.py --steps 2 --warmup 3 --no-cpu"
timeout 120 $CMD > gpurun_out/cap_plain.log 2>&1 || exit 1
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv $CMD > gpurun_out/cap_launches.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:gm_search_kernel -s 65 -c 1 -f -o gpurun_out/prof_r1_sieve $CMD > gpurun_out/cap_sieve.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:gm_dfs_kernel -s 65 -c 1 -f -o gpurun_out/prof_r1_dfs $CMD > gpurun_out/cap_dfs.log 2>&1
ls -la gpurun_out/*_r1*
