"""BASELINE config 4: minimum-overlap sweep 35 / 50 / 75 on the duplicate / contained-read mix at its stated size (seed 4:
2 M x 150 bp at 60x, 30% re-emitted as forward / reverse-complement duplicates, 20% truncated to 100-149 bp).  One bench
line per minimum overlap (device-resident reads, CUDA-event time of the whole hot path, phase breakdown).
    python profiles/run_config4.py [reads] > gpurun_out/r02_config4.jsonl      (parity at this shape: tests/test_gpu_vs_reference_large.py)"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from disco_b200 import gpu, host, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
rs = synth.dup_contained(n, 150, 60.0, seed=4)
packed, lens = host.pack_codes(rs.codes, rs.off, 8)
g = gpu.GpuBuildGraph(0)
for m in (35, 50, 75):
    g.load_reads(packed, lens)
    ms = []
    for it in range(5):
        g.load_reads(packed, lens)
        g.build_graph(m, 4)
        st = g.stats()
        if it >= 2:
            ms.append(st["ms_total"])
    t = float(np.mean(ms))
    print(json.dumps({"metric": "reads/sec overlap-searched", "unit": "reads/s", "value": n / (t / 1000.0), "ms_per_step": t, "n_gpus": 1,
                      "config": {"workload": f"config 4: {n} x 100-150bp, 60x, 30% duplicates, 20% truncated (contained), minOverlap={m}", "min_overlap": m, "reads": n},
                      "phase_ms": {k: float(v) for k, v in st.items() if k.startswith("ms_")},
                      "counters": {k: int(v) for k, v in st.items() if not k.startswith("ms_")}}), flush=True)
g.close()
