#!/usr/bin/env bash
# per-launch device times of one hot-path pass (cold-cache, serialised)
READS=${1:-10000000}; TAG=${2:-launches}
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -s 9 -c 12 --csv --log-file gpurun_out/${TAG}.csv \
    python profiles/profile_step.py $READS 1 > gpurun_out/${TAG}.log 2>&1 < /dev/null
python - <<PY
import csv
rows=[r for r in csv.reader(l for l in open('gpurun_out/${TAG}.csv') if l.startswith('"'))]
h=rows[0]; ki=h.index('Kernel Name'); mi=h.index('Metric Name'); vi=h.index('Metric Value'); ii=h.index('ID')
d={}
for r in rows[1:]:
    d.setdefault(r[ii],{'k':r[ki][:34]})[r[mi].split('.')[0].replace('smsp__','').replace('sm__','')]=r[vi]
for i,v in d.items(): print(i, v)
PY
