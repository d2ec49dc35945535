#!/bin/bash
# profiles/ab_libs.sh MNT "descr..." lib... -- quickbench of library variants under build_ab/ (GPUMOTIF_LIB)
mnt=$1; ds=$2; shift 2
for lib in "$@"; do
  echo "== $lib"
  GPUMOTIF_LIB=$PWD/build_ab/$lib bash profiles/quickbench.sh ab_$lib $mnt $ds
done
