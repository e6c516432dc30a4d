#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== full gpu suite"; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2g9_pytest.log 2>&1; tail -5 gpurun_out/r2g9_pytest.log
for sp in 0 4096 16384 65536 262144; do
  echo "== photons 2e7, small_pass=$sp"; PTL_SMALL_PASS=$sp timeout 300 python scripts/perf_probe.py --species photon --n 20000000 --steps 3 2>&1 | tail -1 | sed 's/.*advance/advance/'
  echo "== electrons 1e4, small_pass=$sp"; PTL_SMALL_PASS=$sp timeout 300 python scripts/perf_probe.py --species electron --n 10000 --steps 4 2>&1 | tail -2 | sed 's/.*advance/advance/'
  echo "== electrons 4e6, small_pass=$sp"; PTL_SMALL_PASS=$sp timeout 300 python scripts/perf_probe.py --species electron --n 4000000 --steps 3 2>&1 | tail -1 | sed 's/.*advance/advance/'
done
echo "== trace photons"; PTL_TRACE=1 timeout 300 python scripts/perf_probe.py --species photon --n 20000000 --steps 2 2>&1 | grep "ptl trace\] step 1" | head -12
echo "== trace electrons 1e4"; PTL_TRACE=1 timeout 300 python scripts/perf_probe.py --species electron --n 10000 --steps 2 2>&1 | grep "ptl trace\] step 1" | head -12
