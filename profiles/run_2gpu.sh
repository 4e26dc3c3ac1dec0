#!/usr/bin/env bash
# 2-GPU check: parity test + a small sharded bench.  usage (gpurun --gpus 2): bash profiles/run_2gpu.sh [reads per gpu]
READS=${1:-2000000}
mkdir -p gpurun_out
nvidia-smi -L
python -m pytest tests/test_multigpu_gpu.py -x -q 2>&1 | tail -5
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 2 --steps 3 --warmup 3 --reads $READS --no-cpu 2>&1 | tail -2 | tee gpurun_out/bench_2gpu.json
python bench.py --gpus 1 --steps 3 --warmup 3 --reads $READS --no-cpu 2>&1 | tail -1 | tee gpurun_out/bench_1gpu_same.json
