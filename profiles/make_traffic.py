"""profiles/traffic_r02.json (read by bench.py -> roofline.traffic) from an ncu launch list:
    python profiles/make_traffic.py gpurun_out/r02x_launches.csv 10000000 "<source label>" > profiles/traffic_r02.json
Takes the LAST launch of every kernel (the warm pass) and sums dram__bytes_read.sum + dram__bytes_write.sum."""
import csv, json, sys

rows = list(csv.reader(open(sys.argv[1])))
reads = int(sys.argv[2])
hdr, per = None, {}
for r in rows:
    if r and r[0] == "ID":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        name = d["Kernel Name"].split("(")[0].replace("void ", "").split("<")[0].strip()
        per.setdefault((int(d["ID"]), name), {})[d["Metric Name"]] = float(d["Metric Value"].replace(",", ""))
last = {}
for (i, name), m in sorted(per.items()):
    last[name] = m
out = {"reads": reads, "source": sys.argv[3] if len(sys.argv) > 3 else sys.argv[1],
       "dram_bytes_per_launch": {k: int(v.get("dram__bytes_read.sum", 0) + v.get("dram__bytes_write.sum", 0)) for k, v in last.items()},
       "ns_per_launch_under_ncu": {k: int(v.get("gpu__time_duration.sum", 0)) for k, v in last.items()},
       "warp_instructions": {k: int(v.get("smsp__inst_executed.sum", 0)) for k, v in last.items()},
       "l2_sector_hit_rate_pct": {k: v.get("lts__t_sector_hit_rate.pct") for k, v in last.items()}}
print(json.dumps(out, indent=1))
