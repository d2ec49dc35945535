#!/bin/bash
# profiles/bisect_sh.sh -- X2 (bitsets + lane state) through the whole GPU suite; X5 = X2 + sequence bytes, X6 = X2 + regex arguments
mkdir -p gpurun_out
GPUMOTIF_LIB=$PWD/build_ab/lib_X2.so timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/x2_pytest.txt 2>&1; tail -3 gpurun_out/x2_pytest.txt
for v in X5 X6; do
  echo "== lib_$v"
  GPUMOTIF_LIB=$PWD/build_ab/lib_$v.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "golden or both_paths" 2>&1 | tail -2
  GPUMOTIF_LIB=$PWD/build_ab/lib_$v.so bash profiles/quickbench.sh bis$v 256 trna pk1 pk_j1+2 descr.trna.general
done
