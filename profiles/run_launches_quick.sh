#!/usr/bin/env bash
# launch list (device time + DRAM bytes per launch) of one warm pass: bash profiles/run_launches_quick.sh [reads] [tag]
READS=${1:-10000000}; TAG=${2:-q}
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,smsp__inst_executed.sum --clock-control none -c 60 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python profiles/profile_step.py $READS 1 > gpurun_out/${TAG}_launch.log 2>&1 < /dev/null
python - <<PY
import csv
rows=list(csv.reader(open("gpurun_out/${TAG}_launches.csv")))
hdr=None; per={}
for r in rows:
    if r and r[0]=="ID": hdr=r; continue
    if hdr and len(r)==len(hdr):
        d=dict(zip(hdr,r)); per.setdefault((int(d["ID"]), d["Kernel Name"].split("(")[0].replace("void ","")[:34]),{})[d["Metric Name"]]=float(d["Metric Value"].replace(",",""))
ids=sorted(per)
half=len(ids)//2
for k in ids[half:]:
    m=per[k]; print("%3d %-34s %8.3f ms  rd %6.2f GB wr %6.2f GB  L2hit %5.1f%%  inst %7.1f M" % (k[0],k[1],m.get("gpu__time_duration.sum",0)/1e6,m.get("dram__bytes_read.sum",0)/1e9,m.get("dram__bytes_write.sum",0)/1e9,m.get("lts__t_sector_hit_rate.pct",0),m.get("smsp__inst_executed.sum",0)/1e6))
PY
