#!/bin/bash
# profiles/prof_dfs.sh TAG MNT descr... -- ncu --set full of the enumeration kernel (gm_dfs_kernel) of each descriptor,
# after a plain run of the same command that exited 0; also prints the survivor counts (GPUMOTIF_DEBUG)
tag=$1; mnt=$2; shift 2
mkdir -p gpurun_out
for d in "$@"; do
  GPUMOTIF_DEBUG=1 python bench.py --descr "$d" --mnt $mnt --steps 1 --warmup 3 --no-cpu > gpurun_out/dfs_${tag}_$d.json 2> gpurun_out/dfs_${tag}_$d.err || continue
  grep "gpumotif: hits" gpurun_out/dfs_${tag}_$d.err | tail -1
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:gm_dfs_kernel -s 3 -c 1 -f \
      -o gpurun_out/prof_${tag}_$d python bench.py --descr "$d" --mnt $mnt --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_${tag}_$d.log 2>&1
done
ls -la gpurun_out/prof_${tag}_*
