#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== full gpu suite"; timeout 2400 python -m pytest tests -m gpu -q -s 2>&1 | grep -E "^\[statistics\]|passed|failed|FAILED|Error" | tail -15 > gpurun_out/r2g10_pytest_summary.log; cat gpurun_out/r2g10_pytest_summary.log
echo "== ncu launch list, photons 2e7"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2g10_photon_launches.csv python scripts/perf_probe.py --species photon --n 20000000 --steps 2 > /dev/null 2>&1; tail -25 gpurun_out/r2g10_photon_launches.csv | cut -c1-200
echo "== sanitizer"; bash scripts/sanitize.sh > /dev/null 2>&1; grep -E "^===|ERROR SUMMARY|RACECHECK SUMMARY|kernel .* small_pass" gpurun_out/r2_sanitizer.log
