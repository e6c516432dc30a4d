#!/bin/bash
# scripts/ab.sh NAME...: device-timed A/B of library variants built by scripts/build_variant.sh (same box, alternating)
cd "$(dirname "$0")/.."
for rep in 1 2; do
for v in "$@"; do
  echo "== $v (rep $rep)"
  PTL_LIB_PATH=$PWD/build/ab/libptl_$v.so python scripts/perf_probe.py --n 4000000 --steps 3 2>&1 | tail -1 | sed 's/.*main_ms/main_ms/'
done; done
