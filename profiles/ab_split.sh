#!/bin/bash
# profiles/ab_split.sh TAG [descr] -- instruction counts / issue utilisation of one (sieve, dfs) launch pair
tag=$1; d=${2:-trna}
GPUMOTIF_PATH=split ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,launch__registers_per_thread,launch__grid_size,launch__block_size,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"gm_search|gm_dfs" -s 6 -c 2 --csv --log-file gpurun_out/ab_$tag.csv python bench.py --descr $d --mnt 64 --steps 1 --warmup 3 --no-cpu > gpurun_out/ab_$tag.log 2>&1
python - gpurun_out/ab_$tag.csv <<'PY'
import csv,sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>5]
h=rows[0]
cur=None
for r in rows[1:]:
    k=r[h.index("Kernel Name")][:34]
    if k!=cur: print(); print(k,end=': '); cur=k
    print(r[h.index("Metric Name")].split('__')[1][:28], r[h.index("Metric Value")],end=' | ')
print()
PY
