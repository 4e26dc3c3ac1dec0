#!/usr/bin/env bash
# ncu passes of one hot-path step (run under gpurun from the repo root):  bash profiles/run_ncu.sh [reads] [tag]
# 1. launch list with per-launch device time (cold-cache, serialised: compare SHARES)
# 2. full capture of the search and reduction kernels of the second pass (first pass = warm-up)
READS=${1:-10000000}
TAG=${2:-r01}
mkdir -p gpurun_out
timeout 150 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 40 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python profiles/profile_step.py $READS 1 > gpurun_out/${TAG}_launch.log 2>&1 < /dev/null
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_edges|k_reduce_mark|k_reduce_emit|k_table_insert|k_contain_uniform" -s 8 -c 8 -f -o gpurun_out/${TAG}_prof \
    python profiles/profile_step.py $READS 1 > gpurun_out/${TAG}_full.log 2>&1 < /dev/null
ls -la gpurun_out
