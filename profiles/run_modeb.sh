#!/usr/bin/env bash
# Mode B (key-sharded table over NVLink peer pointers) on N GPUs: parity test (N = 2), then the bench in both
# partitionings.  usage (gpurun --gpus N): bash profiles/run_modeb.sh N [reads per gpu] [steps] [skip-test] [skip-replicated]
N=${1:-2}; READS=${2:-10000000}; STEPS=${3:-3}
mkdir -p gpurun_out
nvidia-smi -L | head -8
if [ -z "$4" ]; then python -m pytest tests/test_multigpu_gpu.py -x -q -m gpu 2>&1 | tail -5; fi
for part in key-sharded replicated; do
  if [ "$part" = replicated ] && [ -n "$5" ]; then continue; fi
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29677 bench.py --gpus $N --steps $STEPS --warmup 3 --reads $READS --partition $part 2> gpurun_out/bench_${N}gpu_$part.err | tail -1 > gpurun_out/bench_${N}gpu_$part.json
  tail -3 gpurun_out/bench_${N}gpu_$part.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_${N}gpu_$part.json'))
    print('$part', {k:d[k] for k in ('value','ms_per_step','e2e')}); print(d['phase_ms']); print(d['counters'])
except Exception as e:
    print('$part: no result', e)
PY
done
