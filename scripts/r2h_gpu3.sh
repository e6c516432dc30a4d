#!/bin/bash
# session-3 batch 3: layout variants of the warp-private kernel (no compile-time fast selection), e2e shard ramp
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== A/B"; bash scripts/ab.sh base q5nofs q5oohot n5oh n5oh4 n6oh n4oh n5ohl n5h
export PTL_LIB_PATH=$PWD/build/ab/libptl_q5nofs.so
for ramp in 1.0 0.3; do
  echo "== e2e ramp $ramp"; timeout 300 python bench.py --n-per-gpu 20000000 --steps 2 --warmup 2 --no-cpu-baseline --no-secondary --e2e-ramp $ramp 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value %.4g e2e %.4g ms %.1f' % (d['value'], d['e2e']['value'], d['e2e']['ms_per_step']), d['e2e']['worker_phase_ms_last_step'])"
done
export PTL_LIB_PATH=$PWD/build/ab/libptl_n5h.so
echo "== kappa ~ 1: TMA with producer warp"; PTL_KERNEL=tma timeout 200 python scripts/kappa_sweep.py 10000000 one 2>&1 | tail -1 | cut -c1-400
echo "== streaming parity test (both kernels)"; timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "streaming" 2>&1 | tail -3
