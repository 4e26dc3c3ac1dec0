#!/usr/bin/env bash
# N-GPU bench line(s) of round 2 (run under `gpurun --gpus N`): bash profiles/run_ngpu_r02.sh N [tag] [extra bench args]
N=$1; TAG=${2:-r02}; shift 2
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 3 --warmup 3 "$@" \
    > gpurun_out/${TAG}_bench_${N}gpu.json 2> gpurun_out/${TAG}_bench_${N}gpu.err
echo rc=$?
grep -v "^\*\*\*\|OMP_NUM" gpurun_out/${TAG}_bench_${N}gpu.err | tail -6
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_bench_${N}gpu.json").read().strip().splitlines()[-1])
    print("N=$N", "%.1f M reads/s" % (d["value"]/1e6), "%.2f ms" % d["ms_per_step"], "e2e %.1f M" % (d["e2e"]["value"]/1e6), "parity", d["parity_check"]["ok"], "single-GPU same input %s ms" % d["parity_check"].get("single_gpu_ms_same_input"))
    print("   ", {k:round(v,2) for k,v in d["phase_ms"].items()})
    if "config3_shape" in d:
        c=d["config3_shape"]; print("  cfg3 %.1f M reads/s %.2f ms e2e %.1f M parity %s single-GPU same input %.1f ms" % (c["value"]/1e6, c["ms_per_step"], c["e2e"]["value"]/1e6, c["parity_check"]["ok"], c["parity_check"]["single_gpu_ms_same_input"]))
        print("   ", {k:round(v,2) for k,v in c["phase_ms"].items()})
except Exception as e:
    print("ERR", e)
PY
