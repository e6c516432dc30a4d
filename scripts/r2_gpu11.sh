#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== gpu suite"; timeout 2400 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for ov in 1 0; do
echo "== photons 2e7 (overlap=$ov)"; PTL_OVERLAP=$ov timeout 300 python scripts/perf_probe.py --species photon --n 20000000 --steps 3 2>&1 | tail -1 | sed 's/.*advance/advance/'
echo "== electrons 1e4 (overlap=$ov)"; PTL_OVERLAP=$ov timeout 300 python scripts/perf_probe.py --species electron --n 10000 --steps 4 2>&1 | tail -2 | sed 's/.*advance/advance/'
done
echo "== electrons 4e6"; timeout 300 python scripts/perf_probe.py --species electron --n 4000000 --steps 3 2>&1 | tail -1 | sed 's/.*advance/advance/'
echo "== trace photons"; PTL_TRACE=1 timeout 300 python scripts/perf_probe.py --species photon --n 20000000 --steps 2 2>&1 | grep "ptl trace\] step 1" | head -12
echo "== trace electrons 1e4"; PTL_TRACE=1 timeout 300 python scripts/perf_probe.py --species electron --n 10000 --steps 2 2>&1 | grep "ptl trace\] step 1" | head -12
echo "== bench small + secondary full"; timeout 900 python bench.py --n-per-gpu 4000000 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2g11_bench.json 2> gpurun_out/r2g11_bench.err; tail -3 gpurun_out/r2g11_bench.err
