#!/usr/bin/env bash
# N-GPU bench as the driver launches it.  usage (gpurun --gpus N): bash profiles/run_ngpu.sh N [reads per gpu] [steps]
N=${1:-2}; READS=${2:-10000000}; STEPS=${3:-3}
mkdir -p gpurun_out
nvidia-smi -L | head -8; free -g | head -2; nproc
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29677 bench.py --gpus $N --steps $STEPS --warmup 3 --reads $READS 2> gpurun_out/bench_${N}gpu.err | tail -1 | tee gpurun_out/bench_${N}gpu.json
tail -5 gpurun_out/bench_${N}gpu.err
