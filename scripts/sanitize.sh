#!/bin/bash
# compute-sanitizer over every advance-kernel family (verdict r1: "no compute-sanitizer run anywhere").  Logs -> profiles/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
out=gpurun_out/r2_sanitizer.log
: > $out
run() {  # tool kernel small_pass n
  echo "=== compute-sanitizer --tool $1 : kernel $2, small_pass_rows $3, n $4" | tee -a $out
  timeout 1500 compute-sanitizer --tool $1 --print-limit 20 python scripts/sanitize_probe.py --kernel $2 --small-pass $3 --n $4 2>&1 | grep -v "^$" | tail -12 | tee -a $out
}
run memcheck 5 0 50000
run memcheck 3 0 50000
run memcheck 4 0 50000
run memcheck 5 100000000 50000
run racecheck 5 0 20000
run racecheck 3 0 20000
run racecheck 4 0 20000
run synccheck 5 0 20000
run synccheck 3 0 20000
