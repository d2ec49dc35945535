#!/bin/bash
# profiles/win_probe.sh -- bounded probes of the expand8 lane-window build in the enumeration kernel (every command under a short timeout)
mkdir -p gpurun_out
probe() {  # lib, test id
  GPUMOTIF_DEBUG=1 GPUMOTIF_LIB=$PWD/build_ab/$1 timeout 60 python -m pytest "tests/test_gpu_parity.py::test_gpu_matches_reference_golden[$2]" -x -q 2>&1 | grep -E "passed|failed|split path|threads x|Timeout|error" | tail -4
  echo "  ($1 $2 exit ${PIPESTATUS[0]})"
}
echo "== all: pk_nested"; probe lib_win_all.so extra.pk_nested
echo "== lite: pk_nested"; probe lib_win_lite.so extra.pk_nested
echo "== lite: parity suite"; GPUMOTIF_LIB=$PWD/build_ab/lib_win_lite.so timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
echo "== lite: quickbench"; GPUMOTIF_LIB=$PWD/build_ab/lib_win_lite.so timeout 200 bash profiles/quickbench.sh winl 1024 trna descr.trna.general ire
