#!/bin/bash
# session-3 batch 4: current default build against the best measured variant; ncu of a fast and a slow layout of the same code
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== A/B"; bash scripts/ab.sh base q5nofs cur
for v in q5nofs n5oh; do
  echo "== ncu $v"
  PTL_LIB_PATH=$PWD/build/ab/libptl_$v.so timeout 600 ncu --set full --clock-control none -k regex:k_advance_wq -c 1 -f -o gpurun_out/r2h4_$v python scripts/perf_probe.py --n 4000000 --steps 1 > gpurun_out/r2h4_ncu_$v.log 2>&1; tail -1 gpurun_out/r2h4_ncu_$v.log | cut -c1-200
done
