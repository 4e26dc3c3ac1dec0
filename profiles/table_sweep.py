"""Table build alone (begin + phase_table(0), CUDA events inside the library) under the binned build's knobs.
python profiles/table_sweep.py [reads]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from disco_b200 import gpu

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
dev = torch.device("cuda", 0)
d_packed, d_lens = bench.make_packed_on_gpu(n, 2, dev, 8)
torch.cuda.synchronize()


def run(env):
    for k in ("DISCO_BINNED", "DISCO_FILL_BLOCKS", "DISCO_FILL_CHUNK", "DISCO_BIN_SLICE_KB", "DISCO_UNBINNED"):
        os.environ.pop(k, None)
    os.environ.update(env)
    g = gpu.GpuBuildGraph(0)
    g.use_reads_device(d_packed.data_ptr(), d_lens.data_ptr(), n, 8, 150, 150)
    best = 1e9
    for _ in range(4):
        g.begin(50, 4)
        g.phase_table(False)
        g.sync()
        best = min(best, g.stats()["ms_table_all"])
    g.close()
    return best


print("direct", "%.3f ms" % run({"DISCO_BINNED": "0"}), flush=True)
import itertools
for slice_kb, blocks, chunk in itertools.product(("2048", "4096", "16384"), ("4", "8"), ("256", "1024", "4096")):
    env = {"DISCO_BIN_SLICE_KB": slice_kb, "DISCO_FILL_BLOCKS": blocks, "DISCO_FILL_CHUNK": chunk}
    print(env, "%.3f ms" % run(env), flush=True)
