#!/bin/bash
# session-3 batch 10: sanitizer over the kernels this session changed; e2e leg with fewer shards
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
bash scripts/sanitize_final.sh > gpurun_out/r2h10_sanitize_stdout.log 2>&1; grep -E "===|SUMMARY|kernel 5" gpurun_out/r2_final_sanitizer.log | cut -c1-200
run() { tag=$1; shift; timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-secondary --e2e-steps 3 "$@" > gpurun_out/r2h10_$tag.json 2> gpurun_out/r2h10_$tag.err; python - <<P
import json
d=json.loads(open('gpurun_out/r2h10_$tag.json').read().strip().splitlines()[-1]); e=d['e2e']
print('$tag', 'value %.4g e2e %.4g ms %.1f' % (d['value'], e['value'], e['ms_per_step']), e['worker_phase_ms_last_step'])
P
}
run n8 --e2e-shards 8
run n10 --e2e-shards 10
run n10r5 --e2e-shards 10 --e2e-ramp 0.5
