#!/bin/bash
# session-3 batch 1: A/B of the lean scheduler / RBEB trial count / multi-chunk variants, TMA streaming kernel check
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== pytest (streaming path, kernel variants, replay)"; timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "streaming or variants or advance_replay or below_cut" 2>&1 | tail -5
echo "== kappa ~ 1: TMA"; timeout 200 python scripts/kappa_sweep.py 10000000 one 2>&1 | tail -1 | cut -c1-400
echo "== kappa ~ 1: plain"; PTL_KERNEL=notma timeout 200 python scripts/kappa_sweep.py 10000000 one 2>&1 | tail -1 | cut -c1-400
echo "== A/B"; bash scripts/ab.sh base lean2 lean3 lean4 lean5 old4 lean4c2 lean4c3
