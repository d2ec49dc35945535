#!/bin/bash
# profiles/prof_r2_base.sh -- round-2 starting point: GPU parity suite, throughput of the weak-filter
# descriptors and one ncu --set full capture of their dominant kernel (each capture after a plain run
# of the same command that exited 0).
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2base_pytest.txt 2>&1; tail -3 gpurun_out/r2base_pytest.txt
bash profiles/quickbench.sh r2base 64 trna pk1 pk_j1+2 qu+tr descr.trna.general score.1 ire 2>&1 | tee gpurun_out/r2base_quick.txt
for d in pk_j1+2 qu+tr; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:gm_search_kernel -s 3 -c 1 -f \
      -o gpurun_out/prof_r2base_$d python bench.py --descr "$d" --mnt 8 --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_r2base_$d.log 2>&1
done
for d in pk1 descr.trna.general; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:gm_dfs_kernel -s 3 -c 1 -f \
      -o gpurun_out/prof_r2base_$d python bench.py --descr "$d" --mnt 16 --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_r2base_$d.log 2>&1
done
ls -la gpurun_out/prof_r2base_*
