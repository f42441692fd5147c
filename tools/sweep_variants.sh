#!/bin/bash
# Time the simulator kernel of the variants built by tools/build_variants.sh (3.4e7 trials each, one process per
# variant): tools/sweep_variants.sh name1 name2 ...  -> gpurun_out/sweep.log, one JSON line per variant.
ROOT=$(cd "$(dirname "$0")/.." && pwd)
mkdir -p $ROOT/gpurun_out
for v in "$@"; do
  DDM_B200_LIB=$ROOT/build/variants/lib_$v.so python $ROOT/tools/time_sim.py 33554432 2>&1 | tail -1
done > $ROOT/gpurun_out/sweep.log 2>&1
