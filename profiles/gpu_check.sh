#!/bin/bash
# profiles/gpu_check.sh TAG [MNT] [descr...] -- GPU parity suite, then quickbench of the named descriptors
tag=$1; mnt=${2:-64}; shift 2
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.txt 2>&1; tail -15 gpurun_out/${tag}_pytest.txt
bash profiles/quickbench.sh $tag $mnt "$@" 2>&1 | tee gpurun_out/${tag}_quick.txt
