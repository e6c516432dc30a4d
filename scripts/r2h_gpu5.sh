#!/bin/bash
# session-3 batch 5: full GPU test suite on the default build; lepton streaming kernel at 2 / 3 / 4 CTAs per SM; ncu of the TMA-staged one
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== pytest -m gpu"; (time timeout 900 python -m pytest tests -x -q -m gpu) > gpurun_out/r2h5_pytest.log 2>&1; tail -4 gpurun_out/r2h5_pytest.log
for v in s2 cur s4; do
  echo "== kappa ~ 1, lepton streaming kernel, build $v"; PTL_LIB_PATH=$PWD/build/ab/libptl_$v.so timeout 200 python scripts/kappa_sweep.py 10000000 one 2>&1 | tail -1 | cut -c1-420
done
echo "== ncu TMA-staged streaming kernel"
PTL_KERNEL=tma timeout 600 ncu --set full --clock-control none -k regex:k_advance_stream_tma -s 2 -c 1 -f -o gpurun_out/r2h5_stream_tma python scripts/kappa_sweep.py 10000000 one > gpurun_out/r2h5_ncu_tma.log 2>&1; tail -1 gpurun_out/r2h5_ncu_tma.log | cut -c1-300
