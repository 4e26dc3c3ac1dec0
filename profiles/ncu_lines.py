"""Per-source-line instruction / stall-sample shares from an ncu report.
usage: python profiles/ncu_lines.py <report.ncu-rep> <kernel regex> [function-name substring] [top N]"""
import collections, csv, subprocess, sys

rep, kre = sys.argv[1], sys.argv[2]
fsub = sys.argv[3] if len(sys.argv) > 3 else ""
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", f"regex:{kre}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
path = func = None
hdr = None
agg = collections.defaultdict(lambda: [0.0, 0.0, 0.0, ""])   # (file, line) -> inst, samples, thread inst, text
tot_i = tot_s = 0.0
cur_line = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        path = r[1]; hdr = None; continue
    if r[0] == "Function Name":
        func = r[1]; continue
    if r[0] == "Line No":
        hdr = r
        iI, iN, iT = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed")
        continue
    if hdr is None or fsub not in (func or ""):
        continue
    if r[0]:
        cur_line = (path.split("/")[-1], int(r[0]), r[1])
    if len(r) > iT and r[2]:   # a SASS row
        try:
            i, n, t = float(r[iI] or 0), float(r[iN] or 0), float(r[iT] or 0)
        except ValueError:
            continue
        a = agg[cur_line[:2]]
        a[0] += i; a[1] += n; a[2] += t; a[3] = cur_line[2]
        tot_i += i; tot_s += n
print(f"total warp instructions {tot_i:.3e}, samples {tot_s:.0f}")
for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
    eff = a[2] / a[0] if a[0] else 0
    print(f"inst {a[0] / tot_i * 100:5.2f}%  samp {a[1] / max(tot_s, 1) * 100:5.2f}%  thr/inst {eff:4.1f}  {f}:{l:<4} {a[3].strip()[:100]}")
