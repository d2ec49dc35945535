#!/bin/bash
# profiles/r2o_check.sh -- GPU parity suite, default bench, host facts, binary per-batch stats
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2o_pytest.txt 2>&1; tail -5 gpurun_out/r2o_pytest.txt
S=$(date +%s); timeout 900 python bench.py > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err; echo "bench exit $? in $(( $(date +%s) - S )) s"
(lscpu | head -30; nproc; free -g | head -3; nvidia-smi topo -m 2>/dev/null | head -8; cat /sys/bus/pci/devices/*/numa_node 2>/dev/null | sort | uniq -c) > gpurun_out/r2o_host.txt 2>&1
bash profiles/binary_stats.sh 1024 > gpurun_out/r2o_binary.txt 2>&1; tail -40 gpurun_out/r2o_binary.txt
