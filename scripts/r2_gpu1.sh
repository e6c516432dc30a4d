#!/bin/bash
# round 2, GPU call 1: parity of the refactored library + first wq (warp-private) measurement
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2g1_smi.txt 2>&1
echo "== wq quick parity"; timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "variants" > gpurun_out/r2g1_variants.log 2>&1; tail -3 gpurun_out/r2g1_variants.log
for k in bq wq bq wq; do
  echo "== perf $k"; PTL_KERNEL=$k timeout 300 python scripts/perf_probe.py --n 4000000 --steps 3 2>&1 | tail -2 | sed 's/.*advance/advance/'
done | tee gpurun_out/r2g1_perf.log
echo "== full gpu suite"; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2g1_pytest.log 2>&1; tail -5 gpurun_out/r2g1_pytest.log
echo "== ncu wq"; PTL_KERNEL=wq timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_advance_wq -c 1 -o gpurun_out/r2g1_wq_full -f python scripts/perf_probe.py --n 4000000 --steps 1 > gpurun_out/r2g1_ncu.log 2>&1; tail -2 gpurun_out/r2g1_ncu.log
