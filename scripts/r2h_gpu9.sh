#!/bin/bash
# session-3 batch 9: e2e leg, more workers than advance slots (1e8 electrons)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { tag=$1; shift; timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-secondary --e2e-steps 3 "$@" > gpurun_out/r2h9_$tag.json 2> gpurun_out/r2h9_$tag.err; python - <<P
import json
d=json.loads(open('gpurun_out/r2h9_$tag.json').read().strip().splitlines()[-1]); e=d['e2e']
print('$tag', 'value %.4g e2e %.4g ms %.1f' % (d['value'], e['value'], e['ms_per_step']), e['worker_phase_ms_last_step'])
P
}
run w4s3 --e2e-advance-slots 3 --e2e-workers 4 --e2e-shards 16
run w5s3 --e2e-advance-slots 3 --e2e-workers 5 --e2e-shards 15
run w6s4 --e2e-advance-slots 4 --e2e-workers 6 --e2e-shards 18
run w6s3 --e2e-advance-slots 3 --e2e-workers 6 --e2e-shards 18
run w4s2n12 --e2e-advance-slots 2 --e2e-workers 4 --e2e-shards 12
