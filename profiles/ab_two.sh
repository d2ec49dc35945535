#!/bin/bash
# profiles/ab_two.sh -- two-stage sieve on/off for trna at 1024 Mnt, with instruction counts of the sieve kernel
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2h_pytest.txt 2>&1; tail -3 gpurun_out/r2h_pytest.txt
bash profiles/quickbench.sh r2h_two 1024 trna pk_j1+2 descr.trna.general
GPUMOTIF_NO_TWO_STAGE=1 bash profiles/quickbench.sh r2h_one 1024 trna pk_j1+2 descr.trna.general
for v in two one; do
  if [ $v = one ]; then export GPUMOTIF_NO_TWO_STAGE=1; fi
  ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"gm_search|gm_dfs" -s 8 -c 2 --csv --log-file gpurun_out/ab_$v.csv python bench.py --descr trna --mnt 256 --steps 1 --warmup 2 --no-cpu --configs none --no-parity --no-binary > gpurun_out/ab_$v.log 2>&1
  python - gpurun_out/ab_$v.csv <<'PY'
import csv,sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>5]
h=rows[0]
cur=None
for r in rows[1:]:
    k=r[h.index("Kernel Name")][:34]+r[h.index("ID")]
    if k!=cur: print(); print(k,end=': '); cur=k
    print(r[h.index("Metric Name")].split('__')[1][:28], r[h.index("Metric Value")],end=' | ')
print()
PY
done
