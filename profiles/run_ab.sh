#!/usr/bin/env bash
# A/B of an env knob on the 10M bench.  usage: bash profiles/run_ab.sh VAR "v1 v2 ..." [extra bench args]
VAR=$1; VALS=$2; shift 2
mkdir -p gpurun_out
for v in $VALS; do
  env $VAR=$v timeout 200 python bench.py --steps 4 --warmup 3 --no-cpu "$@" 2>&1 | tail -1 > gpurun_out/ab_$v.json
  python - <<PY
import json
d=json.load(open('gpurun_out/ab_$v.json')); p=d['phase_ms']
print('$VAR=$v: %.1f M reads/s step %.2f ms e2e %.2f | probe %.2f verify %.2f exact %.2f | mark %.2f emit %.2f' % (d['value']/1e6, d['ms_per_step'], d['e2e']['ms_per_step'], p['ms_edges_probe'], p['ms_edges_verify'], p['ms_edges_exact'], p['ms_mark_kernel'], p['ms_emit_kernel']))
PY
done
