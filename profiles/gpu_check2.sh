#!/bin/bash
# profiles/gpu_check2.sh TAG -- GPU parity suite, then the weak-filter descriptors at 64 and 1024 Mnt
tag=$1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.txt 2>&1; tail -5 gpurun_out/${tag}_pytest.txt
bash profiles/quickbench.sh ${tag}_64 64 pk1 pk_j1+2 qu+tr descr.trna.general 2>&1 | tee gpurun_out/${tag}_quick.txt
bash profiles/quickbench.sh ${tag}_1024 1024 pk1 pk_j1+2 qu+tr descr.trna.general 2>&1 | tee -a gpurun_out/${tag}_quick.txt
