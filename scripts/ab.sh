#!/bin/bash
# scripts/ab.sh NAME...: device-timed A/B of library variants built by scripts/build_variant.sh (same box, alternating).
# Prints per variant the main-pass time of the last step plus substeps / births (these must be identical across variants
# that claim identical results).  N (rows) and REPS from the environment.
cd "$(dirname "$0")/.."
N=${N:-4000000}; REPS=${REPS:-2}
for rep in $(seq 1 $REPS); do
for v in "$@"; do
  echo -n "== $v (rep $rep) "
  PTL_LIB_PATH=$PWD/build/ab/libptl_$v.so python scripts/perf_probe.py --n $N --steps 3 2>&1 | tail -1 | sed -E 's/.*(substeps=[0-9]+).*(births=[0-9]+).*(main_ms=[0-9.]+).*/\1 \2 \3/'
done; done
