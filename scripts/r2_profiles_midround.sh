#!/bin/bash
# round-2 evidence run: ncu launch list of the bench command, full captures of the dominant kernels, kappa sweep
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B="python bench.py --n-per-gpu 20000000 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary"
echo "== launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_bench_launches.csv $B > gpurun_out/r2_bench_under_ncu.log 2>&1; tail -2 gpurun_out/r2_bench_under_ncu.log
echo "== full capture: electron first pass"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_advance_wq -s 3 -c 1 -o gpurun_out/r2_wq_full -f $B > gpurun_out/r2_ncu_full.log 2>&1; tail -1 gpurun_out/r2_ncu_full.log
echo "== full capture: photon streaming"; timeout 600 ncu --set full --clock-control none -k regex:k_advance_stream -s 2 -c 1 -o gpurun_out/r2_photon_stream_full -f python scripts/perf_probe.py --species photon --n 20000000 --steps 3 > gpurun_out/r2_ncu_photon.log 2>&1; tail -1 gpurun_out/r2_ncu_photon.log
echo "== full capture: electron streaming (kappa ~ 1)"; timeout 600 ncu --set full --clock-control none -k regex:k_advance_stream -s 2 -c 1 -o gpurun_out/r2_electron_stream_full -f python scripts/kappa_sweep.py 10000000 one > gpurun_out/r2_ncu_estream.log 2>&1; tail -1 gpurun_out/r2_ncu_estream.log
echo "== kappa sweep"; timeout 900 python scripts/kappa_sweep.py 10000000 > gpurun_out/r2_kappa_sweep.jsonl 2> gpurun_out/r2_kappa.err; cat gpurun_out/r2_kappa_sweep.jsonl | cut -c1-260
