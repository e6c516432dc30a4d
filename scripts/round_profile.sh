#!/bin/bash
# One GPU-box batch: parity tests, the full default bench, the ncu launch list of a shorter bench, one ncu --set full capture of
# the dominant kernel, the kappa sweep, and the LXCat / photon side measurements.  Outputs under gpurun_out/$TAG_*.
TAG=${1:-r1s2}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) > gpurun_out/${TAG}_pytest.log 2>&1; tail -3 gpurun_out/${TAG}_pytest.log
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; cat gpurun_out/${TAG}_bench.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err; cat gpurun_out/${TAG}_bench_reference.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_bench_launches.csv \
    python bench.py --n-per-gpu 20000000 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_advance_bq -c 1 -f -o gpurun_out/${TAG}_bq_full \
    python bench.py --n-per-gpu 20000000 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_ncu_full.log 2>&1
python scripts/kappa_sweep.py > gpurun_out/${TAG}_kappa_sweep.jsonl 2> gpurun_out/${TAG}_kappa.err; cat gpurun_out/${TAG}_kappa_sweep.jsonl | cut -c1-200
for np in 7 64; do python scripts/perf_probe.py --species slow --n 10000000 --steps 3 --lx-procs $np 2>&1 | tail -1; done | tee gpurun_out/${TAG}_lxcat.log
python scripts/perf_probe.py --species photon --n 20000000 --steps 3 2>&1 | tail -2 | tee gpurun_out/${TAG}_photon.log
python scripts/mixed_probe.py --ne 50000000 --ng 50000000 --npos 1000000 --steps 3 2>&1 | tail -1 | tee gpurun_out/${TAG}_mixed.jsonl
ls -la gpurun_out | tail -20
