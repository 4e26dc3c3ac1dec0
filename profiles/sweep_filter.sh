#!/usr/bin/env bash
# filter-size sweep of the 10M-read bench (tuning aid)
for lg in 0 29 28 27 26 25; do
  echo "== DISCO_FILTER_LOG2=$lg"
  DISCO_FILTER_LOG2=$lg python bench.py --steps 3 --warmup 3 --no-cpu "$@" 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], {k:round(v,2) for k,v in d['phase_ms'].items()}, d['counters']['buckets_edges'])"
done
