#!/usr/bin/env bash
# gpurun helper: GPU tests then the 10M-read bench (no CPU leg).  usage: bash profiles/run_bench.sh [extra bench args]
# (every step under its own timeout: a hung kernel must not eat the GPU budget)
timeout 400 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu "$@" 2>&1 | tail -1 > gpurun_out/bench_last.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_last.json'))
print({k:d[k] for k in ('value','ms_per_step','e2e','roofline','phase_ms')})
print(d['counters'])
PY
