#!/usr/bin/env bash
# A/B of env knobs on the 10M bench, several VAR=value sets.  usage: bash profiles/run_ab2.sh "A=1 B=2" "A=0" ...   (extra bench args via BENCH_ARGS)
mkdir -p gpurun_out
i=0
for kv in "$@"; do
  i=$((i+1))
  env $kv timeout 200 python bench.py --steps 4 --warmup 3 --no-cpu --no-config3 $BENCH_ARGS 2>&1 | tail -1 > gpurun_out/ab2_$i.json
  KV="$kv" python - <<PY
import json, os
try:
    d=json.load(open('gpurun_out/ab2_$i.json')); p=d['phase_ms']; c=d['counters']
    print('%-40s %.1f M reads/s step %.2f ms e2e %.2f | tab %.2f+%.2f cont %.2f | probe %.2f verify %.2f exact %.2f | mark %.2f emit %.2f | slow %d buckets %d' % (os.environ['KV'], d['value']/1e6, d['ms_per_step'], d['e2e']['ms_per_step'], p['ms_table_all'], p['ms_table_nc'], p['ms_contained'], p['ms_edges_probe'], p['ms_edges_verify'], p['ms_edges_exact'], p['ms_mark_kernel'], p['ms_emit_kernel'], c['slow_path_reads'], c['buckets_edges']))
except Exception as e:
    print(os.environ['KV'], 'FAILED', e, open('gpurun_out/ab2_$i.json').read()[-300:])
PY
done
