"""The one full-size CPU run of BASELINE config 2: the reference's OpenMP buildG (oracle/_ref/buildG, all host cores) on the
SAME 10 M x 150 bp FASTA the GPU path then builds its graph from, and a file-level comparison of the two results.
Run under gpurun from the repo root:  python profiles/ref_full_config2.py [reads] > gpurun_out/r02_ref_full_config2.json"""
import json, os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from disco_b200 import synth
from disco_b200.buildgraph import BuildGraph
from oracle import oracle

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
cores = os.cpu_count() or 1
t0 = time.time()
rs = synth.single_genome(n, 150, 30.0, seed=2)
d = tempfile.mkdtemp(prefix="disco_full_")
fa = os.path.join(d, "reads.fa")
rs.write_fasta(fa)
t_gen = time.time() - t0
t0 = time.time()
ref = oracle.run_ref([fa], os.path.join(d, "ref", "o"), 50, threads=cores, mem_gb=200, timeout=3000)
t_ref_wall = time.time() - t0
t0 = time.time()
bg = BuildGraph(min_overlap=50, device=0)
bg.add_file(fa)
res = bg.run()
t_gpu_wall = time.time() - t0
mine = sorted(bg.edge_lines())
rows = bg.crow_lines()
st = res.stats
out = {
    "what": "BASELINE config 2 at full size on the GPU box's host cores, reference vs GPU on the same FASTA",
    "reads": n, "cores": cores, "reference_returncode": ref["returncode"],
    "reference_seconds": {k: ref["times"].get(k) for k in ("insertDataset", "buildOverlapGraphFromHashTable")},
    "reference_wall_seconds": t_ref_wall,
    "reference_reads_per_s": n / (ref["times"].get("insertDataset", 0.0) + ref["times"]["buildOverlapGraphFromHashTable"]),
    "gpu_device_ms": float(st["ms_total"]), "gpu_reads_per_s_device": n / (st["ms_total"] / 1000.0),
    "gpu_wall_seconds_incl_parse": t_gpu_wall,
    "edges_gpu": len(mine), "edges_reference": len(ref["edges"]), "edges_equal": mine == ref["edges"],
    "contained_gpu": len(rows), "contained_reference": len(ref["contained_set"]),
    "contained_set_equal": set(int(x.split("\t")[0]) for x in rows) == ref["contained_set"],
    "cap_fired": int(st["cap_fired"]), "multi_overlap_pairs": int(st["multi_overlap_pairs"]), "one_sided_edges": int(st["one_sided_edges"]),
    "seconds_generating_input": t_gen,
}
print(json.dumps(out))
