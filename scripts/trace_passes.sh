#!/bin/bash
# per-pass wall times of the bench (PTL_TRACE) next to the per-step host times, several runs: where does step-to-step variance come from?
cd "$(dirname "$0")/.."
for i in 1 2 3 4; do
  PTL_TRACE=1 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/tp_$i.json 2> gpurun_out/tp_$i.err
  python - <<PY
import json,re
d=json.load(open("gpurun_out/tp_$i.json")); print("run $i", round(d["value"]/1e6,1), "M/s", round(d["ms_per_step"],1), "ms/step main", round(d["roofline"]["kernel_ms_per_launch"],1), d["host_ms_per_step"])
steps={}
for l in open("gpurun_out/tp_$i.err"):
    m=re.match(r"\[ptl trace\] step (\d+) pass (\d+): rows (\d+) substeps (\d+)\s+([\d.]+) ms",l)
    if m: steps.setdefault(int(m.group(1)),[]).append((int(m.group(2)),int(m.group(3)),float(m.group(5))))
for s in sorted(steps):
    if s>=3: print("  step",s," ".join(f"p{p}:{ms:.1f}" for p,r,ms in steps[s]), "sum %.1f"%sum(ms for _,_,ms in steps[s]))
PY
done
