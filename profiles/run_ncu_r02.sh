#!/usr/bin/env bash
# round 2: ncu passes of one hot-path step at 10M reads (run under gpurun from the repo root): bash profiles/run_ncu_r02.sh [reads] [tag]
# 1. launch list with per-launch device time + DRAM bytes (cold-cache, serialised: compare SHARES)
# 2. full capture of the hot kernels of the second pass (first pass = warm-up)
READS=${1:-10000000}
TAG=${2:-r02}
mkdir -p gpurun_out
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,smsp__inst_executed.sum --clock-control none -c 60 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python profiles/profile_step.py $READS 1 > gpurun_out/${TAG}_launch.log 2>&1 < /dev/null
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_probe_flat|k_verify_flat|k_edges|k_reduce_mark|k_reduce_emit|k_table_bin|k_table_fill|k_contain_uniform" -s 9 -c 9 -f -o gpurun_out/${TAG}_prof \
    python profiles/profile_step.py $READS 1 > gpurun_out/${TAG}_full.log 2>&1 < /dev/null
ncu -i gpurun_out/${TAG}_prof.ncu-rep --page raw --csv > gpurun_out/${TAG}_ncu_full.csv 2>/dev/null
for k in probe_flat verify_flat reduce_mark reduce_emit; do python profiles/ncu_lines.py gpurun_out/${TAG}_prof.ncu-rep $k > gpurun_out/${TAG}_ncu_lines_$k.txt 2>&1; done
ls -la gpurun_out | tail -12
