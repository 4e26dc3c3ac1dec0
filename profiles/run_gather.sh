#!/usr/bin/env bash
mkdir -p gpurun_out
./profiles/gather_bench 4 0 | tee gpurun_out/gather_default.txt
./profiles/gather_bench 4 32 | tee gpurun_out/gather_g32.txt
timeout 300 ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum,lts__t_sectors_op_read.sum --clock-control none --csv --log-file gpurun_out/gather_ncu.csv ./profiles/gather_bench 4 0 > /dev/null 2>&1 < /dev/null
timeout 300 ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum,lts__t_sectors_op_read.sum --clock-control none --csv --log-file gpurun_out/gather_ncu_g32.csv ./profiles/gather_bench 4 32 > /dev/null 2>&1 < /dev/null
