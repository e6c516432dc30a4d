#!/bin/bash
# compute-sanitizer over the kernels this session changed: the warp-private kernel (five RBEB trials per unit, Philox call in
# the STEP unit) and the lepton streaming kernels (plain at two CTAs per SM, TMA-staged with mbarriers).  Log -> profiles/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
out=gpurun_out/r2_final_sanitizer.log
: > $out
run() {  # tool n extra-args...
  tool=$1; n=$2; shift 2
  echo "=== compute-sanitizer --tool $tool : n $n $*" | tee -a $out
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_probe.py --n $n "$@" 2>&1 | grep -v "^$" | tail -8 | tee -a $out
}
run memcheck 50000 --kernel 5
run racecheck 20000 --kernel 5
run synccheck 20000 --kernel 5
for tma in 0 1; do
  run memcheck 30000 --kernel 5 --steps 3 --dt-scale 0.00048828125 --stream-tma $tma
  run racecheck 20000 --kernel 5 --steps 3 --dt-scale 0.00048828125 --stream-tma $tma
  run synccheck 20000 --kernel 5 --steps 3 --dt-scale 0.00048828125 --stream-tma $tma
done
