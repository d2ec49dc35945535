#!/bin/bash
# profiles/capture_r2b.sh TAG -- ncu of the headline command only (trna, 1024 Mnt): launch list, --set full of the
# bench-size filter-kernel and enumeration-kernel launches (launch 65: the first scan of a context runs 64 default segments)
tag=${1:-r2b}
set -x
mkdir -p gpurun_out
HEAD="python bench.py --steps 2 --warmup 3 --no-cpu --configs none --no-parity --no-binary --upload chars"
timeout 200 $HEAD > gpurun_out/${tag}_plain.log 2>&1 || exit 1
tail -c 600 gpurun_out/${tag}_plain.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${tag}.csv $HEAD > gpurun_out/${tag}_launches.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gm_filter_kernel -s 65 -c 1 -f -o gpurun_out/prof_${tag}_sieve $HEAD > gpurun_out/${tag}_sieve.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gm_dfs_kernel -s 65 -c 1 -f -o gpurun_out/prof_${tag}_dfs $HEAD > gpurun_out/${tag}_dfs.log 2>&1
for r in gpurun_out/prof_${tag}_*.ncu-rep; do
  b=$(basename $r .ncu-rep)
  python profiles/summarize.py full $r > gpurun_out/${b#prof_}.txt 2>&1
done
python profiles/summarize.py json gpurun_out/prof_${tag}_sieve.ncu-rep trna 1024 > gpurun_out/${tag}_sieve_kernel.json
python profiles/summarize.py launches gpurun_out/launches_${tag}.csv > gpurun_out/${tag}_launches.txt
rm -f gpurun_out/prof_${tag}_*.ncu-rep
