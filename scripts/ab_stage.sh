#!/bin/bash
# A/B of library builds: scripts/ab_stage.sh F reps lib1.so lib2.so ...   ("-" = the in-tree build)
F=$1; R=$2; shift 2
for lib in "$@"; do
  if [ "$lib" = "-" ]; then unset FBPR_B200_LIB; else export FBPR_B200_LIB=$PWD/$lib; fi
  echo "== $lib: $(python scripts/stage_times.py $F $R 2>&1 | tail -1)"
done
