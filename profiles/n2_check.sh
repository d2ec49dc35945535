#!/bin/bash
# profiles/n2_check.sh -- the N>1 launch of the bench as the driver does it, and rnamotif_gpu over two GPUs
mkdir -p gpurun_out
S=$(date +%s)
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2_n2_bench.json 2> gpurun_out/r2_n2_bench.err
echo "bench N=2 exit $? in $(( $(date +%s) - S )) s"; tail -3 gpurun_out/r2_n2_bench.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/r2_n2_ref.json 2> gpurun_out/r2_n2_ref.err
echo "reference arm N=2 exit $?"; cat gpurun_out/r2_n2_ref.json | cut -c1-300
python - <<PY
import sys; sys.path.insert(0, '.')
import bench
bench.write_fasta('/dev/shm/syn.fastn', 2048, 1_000_000, 1001)
PY
cd oracle/_ref/data/test
for dev in 0 0,1; do
  S=$(date +%s.%N)
  EFNDATA=../efndata GPUMOTIF_STATS=1 GPUMOTIF_DEVICES=$dev $OLDPWD/rnamotif_b200/host/_build/rnamotif_gpu -descr trna.descr /dev/shm/syn.fastn > /dev/shm/out_$dev.txt 2> /dev/shm/err.txt
  echo "devices $dev: $(python -c "import time; print(round(time.time()-$S,2))") s"; grep "wall" /dev/shm/err.txt
done
cmp /dev/shm/out_0.txt /dev/shm/out_0,1.txt && echo "stdout identical on 1 and 2 GPUs ($(wc -c < /dev/shm/out_0.txt) bytes)"
rm -f /dev/shm/syn.fastn /dev/shm/out_*.txt /dev/shm/err.txt
