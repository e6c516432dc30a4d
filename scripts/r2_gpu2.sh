#!/bin/bash
# round 2, GPU call 2: CTA-synchronous warp-private kernel variants (A/B against bq on the same box)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in "$@"; do
  echo "== parity $v"; PTL_LIB_PATH=$PWD/build/ab/libptl_$v.so timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "variants" 2>&1 | tail -1
done
for rep in 1 2; do
for v in "$@"; do
  echo "== $v (rep $rep)"
  PTL_LIB_PATH=$PWD/build/ab/libptl_$v.so timeout 300 python scripts/perf_probe.py --n 4000000 --steps 3 2>&1 | tail -1 | sed 's/.*main_ms/main_ms/'
done; done
