#!/bin/bash
# profiles/ab_win.sh -- enumeration kernel: lane windows built eight nucleotides at a time (lib_new) vs one (lib_old)
mkdir -p gpurun_out
GPUMOTIF_LIB=$PWD/build_ab/lib_new.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/win_pytest.txt 2>&1; tail -3 gpurun_out/win_pytest.txt
bash profiles/ab_libs.sh 1024 "trna pk1 pk_j1+2 descr.trna.general qu+tr" lib_old.so lib_new.so
