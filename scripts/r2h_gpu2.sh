#!/bin/bash
# session-3 batch 2: RBEB trial count with the full-recount scheduler, compile-time fast selection, outlined OTHER unit, hot-function placement
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PTL_LIB_PATH=$PWD/build/ab/libptl_q5.so
echo "== kappa ~ 1: plain (inline cheb3 setr)"; timeout 200 python scripts/kappa_sweep.py 10000000 one 2>&1 | tail -1 | cut -c1-400
echo "== kappa ~ 1: TMA"; PTL_KERNEL=tma timeout 200 python scripts/kappa_sweep.py 10000000 one 2>&1 | tail -1 | cut -c1-400
unset PTL_LIB_PATH
echo "== A/B"; bash scripts/ab.sh base old4 q4 q5 q6 q8 q5nofs q5oo q5oohot
