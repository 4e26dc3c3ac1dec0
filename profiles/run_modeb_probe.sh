#!/usr/bin/env bash
# Mode B on N GPUs: symmetric memory (VMM, 2 MB pages) against legacy CUDA IPC mappings of cudaMalloc buffers.
# usage (gpurun --gpus N): bash profiles/run_modeb_probe.sh N [reads per gpu] ["1 0" = DISCO_SYMM variants]
N=${1:-2}; READS=${2:-10000000}; VARIANTS=${3:-"1 0"}
mkdir -p gpurun_out
run() { # reads symm
  DISCO_SYMM=$2 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29677 bench.py --gpus $N --steps 3 --warmup 3 --reads $1 --partition key-sharded 2> gpurun_out/mb.err | tail -1 > gpurun_out/bench_${N}gpu_key-sharded_symm$2.json
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_${N}gpu_key-sharded_symm$2.json')); p=d['phase_ms']
    print('N=$N reads/gpu $1 symm=$2: %.1f M reads/s, step %.1f ms (e2e %.1f) | tables %.1f+%.1f contain %.1f | probe %.1f verify %.1f | mark kernel %.1f (phase %.1f) emit kernel %.1f (phase %.1f)' % (d['value']/1e6, d['ms_per_step'], d['e2e']['ms_per_step'], p['ms_table_all'], p['ms_table_nc'], p['ms_contained'], p['ms_edges_probe'], p['ms_edges_verify'], p['ms_mark_kernel'], p['ms_mark'], p['ms_emit_kernel'], p['ms_emit']))
except Exception as e:
    print('no result', e); print(open('gpurun_out/mb.err').read()[-3000:])
PY
}
if [ "$N" = 2 ] && [ -z "$4" ]; then python -m pytest tests/test_multigpu_gpu.py -x -q -m gpu 2>&1 | tail -5; fi
for v in $VARIANTS; do run $READS $v; done
