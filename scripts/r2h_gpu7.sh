#!/bin/bash
# session-3 batch 7: STEP unit with out-of-line Philox / log, one-reciprocal RBEB trial; photon streaming kernel check; e2e shard timeline
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== A/B"; bash scripts/ab.sh base fin pc lc pclc rcp
export PTL_LIB_PATH=$PWD/build/ab/libptl_fin.so
echo "== photon streaming (fin)"; timeout 200 python scripts/perf_probe.py --species photon --n 20000000 --steps 3 2>&1 | tail -1 | cut -c1-300
echo "== RBEB event replay with the one-reciprocal trial"; PTL_LIB_PATH=$PWD/build/ab/libptl_rcp.so timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "collide_events or advance_replay or mixed_population or variants" 2>&1 | tail -3
echo "== e2e timeline (1e8 electrons)"; timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-secondary --e2e-steps 2 > gpurun_out/r2h7_e2e.json 2> gpurun_out/r2h7_e2e.err; cut -c1-300 gpurun_out/r2h7_e2e.json
