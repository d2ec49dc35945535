#!/bin/bash
# profiles/prof_descr.sh -- per-descriptor throughput + one ncu --set full capture of the
# search kernel each (run on the GPU box: gpurun -- 'bash profiles/prof_descr.sh TAG descr...').
# Each capture is taken after the same command has exited 0 without ncu.
tag=$1; shift
for d in "$@"; do
  python bench.py --descr "$d" --mnt 32 --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_${tag}_$d.json 2> gpurun_out/bench_${tag}_$d.err || continue
  python bench.py --descr "$d" --mnt 4 --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1 || continue
  ncu --set full --clock-control none --import-source on -k regex:gm_search_kernel -s 3 -c 1 -f \
      -o gpurun_out/prof_${tag}_$d python bench.py --descr "$d" --mnt 4 --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_${tag}_$d.log 2>&1
done
grep -h -o '"value": [0-9.]*, "unit"\|"workload": "[^ ]*' gpurun_out/bench_${tag}_*.json
