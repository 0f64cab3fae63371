#!/bin/bash
# Sweep of the library trace-kernel variants (HIPRT_B200_TRACE_VARIANT) on the bench workloads.
# usage: tools/sweep_trace.sh <tag> "<variants>" "<workloads>"
tag=$1; variants=${2:-"1 10 11 12"}; workloads=${3:-"c2_sphere_1M c3_terrain_10M"}
out=gpurun_out/$tag; mkdir -p $out
for w in $workloads; do
  for v in $variants; do
    HIPRT_B200_TRACE_VARIANT=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload $w > $out/sweep_${w}_v$v.json 2> $out/sweep_${w}_v$v.err
    python - <<PY
import json
try:
    d = json.loads(open("$out/sweep_${w}_v$v.json").read().strip().splitlines()[-1])
    print("$w variant $v: value", d["value"], "Mrays/s  ms/step", d["ms_per_step"], " primary GB/s", d["roofline"]["achieved"], " e2e", d["e2e"]["value"])
except Exception as e:
    print("$w variant $v: FAILED", e, open("$out/sweep_${w}_v$v.err").read()[-600:])
PY
  done
done | tee $out/sweep.txt
