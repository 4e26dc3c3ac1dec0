#!/usr/bin/env bash
# compute-sanitizer passes over a small end-to-end run (all kernels incl. the exact path and variable lengths)
mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import sys; sys.path.insert(0, '.')
from disco_b200 import synth
from disco_b200.buildgraph import BuildGraph
for name, rs, m in (("single", synth.single_genome(3000, 150, 30.0, seed=1), 50),
                    ("varlen", synth.dup_contained(3000, 150, 60.0, seed=4), 35),
                    ("repeats", synth.repeats(1500, 150, seed=11), 50)):
    bg = BuildGraph(min_overlap=m); bg.add_records(rs.strings()); res = bg.run()
    print(name, res.n, len(res.crows), len(res.edges), res.stats["cap_fired"], res.stats["slow_path_reads"]); bg.close()
PY
for tool in memcheck racecheck synccheck; do
  echo "== $tool"
  timeout 600 compute-sanitizer --tool $tool --print-limit 5 python /tmp/san.py 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Error|error|Hazard|single|varlen|repeats" | head -12
done 2>&1 | tee gpurun_out/sanitizer.txt
