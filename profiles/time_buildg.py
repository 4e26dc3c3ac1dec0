"""Wall time of the buildG executable on BASELINE config 2's input (10 M x 150 bp FASTA, -t 16): the stage timings it
prints (Dataset = parse + filter + pack, the GPU stage, the file writers) and the total.
Run under gpurun from the repo root:  python profiles/time_buildg.py [reads] [shards]"""
import os, re, subprocess, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from disco_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
shards = int(sys.argv[2]) if len(sys.argv) > 2 else (os.cpu_count() or 1)
rs = synth.single_genome(n, 150, 30.0, seed=2)
d = tempfile.mkdtemp(prefix="disco_buildg_")
fa = os.path.join(d, "reads.fa")
rs.write_fasta(fa)
cfg = os.path.join(d, "disco.cfg")
open(cfg, "w").write("MinOverlap4BuildGraph = 50\n")
exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "disco_b200", "bin", "buildG")
for it in range(2):
    pre = os.path.join(d, f"run{it}", "g")
    os.makedirs(os.path.dirname(pre))
    t0 = time.time()
    r = subprocess.run([exe, "-se", fa, "-f", pre, "-p", cfg, "-t", str(shards)], capture_output=True, text=True)
    wall = time.time() - t0
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    stages = re.findall(r"Function (\w+)\(\) finished in ([0-9.e+-]+) Seconds", r.stdout)
    size = sum(os.path.getsize(os.path.join(os.path.dirname(pre), f)) for f in os.listdir(os.path.dirname(pre)))
    print([l for l in r.stdout.splitlines() if l.startswith("GPU stage")][:1])
    print(f"run {it}: wall {wall:.2f} s, {n} reads, {shards} shards, {size / 1e6:.0f} MB written |", ", ".join(f"{a} {float(b):.3f}" for a, b in stages), flush=True)
