#!/usr/bin/env bash
# timing ablations of the edge-search kernel (outputs are wrong by construction; only ms_edges_kernel matters)
for d in 0 2 4 1; do
  echo "== DISCO_DBG=$d"
  DISCO_DBG=$d python bench.py --steps 2 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('edges kernel ms', round(d['phase_ms']['ms_edges_kernel'],2), 'mark', round(d['phase_ms']['ms_mark'],2))"
done
