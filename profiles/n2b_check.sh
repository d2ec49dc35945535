#!/bin/bash
# profiles/n2b_check.sh -- the N=2 launch of the bench as the driver does it (per-rank NUMA binding, host-pack threads per rank)
mkdir -p gpurun_out
S=$(date +%s)
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2_n2_bench.json 2> gpurun_out/r2_n2_bench.err
echo "bench N=2 exit $? in $(( $(date +%s) - S )) s"; tail -3 gpurun_out/r2_n2_bench.err
nproc; nvidia-smi topo -m | head -6; cat /sys/bus/pci/devices/*/numa_node | sort | uniq -c
