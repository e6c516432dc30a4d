#!/bin/bash
# One GPU-box batch on the final tree of a round: streaming kernels, full GPU test suite, the driver's bench commands, ncu evidence
# (outputs gpurun_out/r2f_*; the summaries copied to profiles/r2_final_* come from here)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== kappa ~ 1: plain"; timeout 200 python scripts/kappa_sweep.py 10000000 one 2>&1 | tail -1 | cut -c1-420
echo "== kappa ~ 1: TMA (2 CTAs/SM, 4 stages)"; PTL_KERNEL=tma timeout 200 python scripts/kappa_sweep.py 10000000 one 2>&1 | tail -1 | cut -c1-420
echo "== pytest -m gpu"; (time timeout 900 python -m pytest tests -x -q -m gpu) > gpurun_out/r2f_pytest.log 2>&1; grep -E "passed|failed" gpurun_out/r2f_pytest.log | tail -2
echo "== bench (default flags)"; python bench.py > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; cut -c1-600 gpurun_out/r2f_bench.json
echo "== bench --impl reference"; python bench.py --impl reference > gpurun_out/r2f_bench_reference.json 2>> gpurun_out/r2f_bench.err; cut -c1-300 gpurun_out/r2f_bench_reference.json
B="python bench.py --n-per-gpu 20000000 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary"
echo "== launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2f_bench_launches.csv $B > gpurun_out/r2f_bench_under_ncu.log 2>&1; tail -1 gpurun_out/r2f_bench_under_ncu.log | cut -c1-200
echo "== full capture: electron first pass"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_advance_wq -s 3 -c 1 -o gpurun_out/r2f_wq_full -f $B > gpurun_out/r2f_ncu_full.log 2>&1; tail -1 gpurun_out/r2f_ncu_full.log | cut -c1-200
echo "== full capture: electron streaming (kappa ~ 1)"; timeout 600 ncu --set full --clock-control none -k regex:k_advance_stream -s 2 -c 1 -o gpurun_out/r2f_electron_stream_full -f python scripts/kappa_sweep.py 10000000 one > gpurun_out/r2f_ncu_estream.log 2>&1; tail -1 gpurun_out/r2f_ncu_estream.log | cut -c1-200
echo "== kappa sweep"; timeout 900 python scripts/kappa_sweep.py 10000000 > gpurun_out/r2f_kappa_sweep.jsonl 2> gpurun_out/r2f_kappa.err; cut -c1-200 gpurun_out/r2f_kappa_sweep.jsonl
echo "== full capture: photon streaming"; timeout 600 ncu --set full --clock-control none -k regex:k_advance_stream -s 2 -c 1 -o gpurun_out/r2f_photon_stream_full -f python scripts/perf_probe.py --species photon --n 20000000 --steps 3 > gpurun_out/r2f_ncu_photon.log 2>&1; tail -1 gpurun_out/r2f_ncu_photon.log | cut -c1-200
