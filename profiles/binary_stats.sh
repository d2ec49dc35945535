#!/bin/bash
# profiles/binary_stats.sh MNT -- rnamotif_gpu over a synthetic FASTA file on tmpfs with per-batch timings
mnt=${1:-1024}
python - <<PY
import sys; sys.path.insert(0, '.')
import bench
bench.write_fasta('/dev/shm/syn.fastn', $mnt, 1_000_000, 1001)
PY
cd oracle/_ref/data/test
for i in 1 2; do
  /usr/bin/env time -f "wall %e s" true 2>/dev/null
  S=$(date +%s.%N)
  EFNDATA=../efndata GPUMOTIF_STATS=1 $OLDPWD/rnamotif_b200/host/_build/rnamotif_gpu -descr trna.descr /dev/shm/syn.fastn > /dev/shm/out.txt 2> /dev/shm/err.txt
  E=$(date +%s.%N)
  echo "run $i: $(echo "$E - $S" | bc -l 2>/dev/null || python -c "print($E-$S)") s"
  grep -v "^trna.descr" /dev/shm/err.txt | tail -14
done
rm -f /dev/shm/syn.fastn /dev/shm/out.txt /dev/shm/err.txt
