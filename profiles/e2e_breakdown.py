"""Where the end-to-end step (bench.py `e2e`) spends its time beyond the device path: host wall clock around each stage of
one C-ABI round trip with pinned host buffers, 10 M x 150 bp (run under gpurun: python profiles/e2e_breakdown.py [reads])."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from disco_b200 import gpu  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
d_packed, d_lens = bench.make_packed_on_gpu(n, 2, dev, 8)
hwpr = 5
h_packed = torch.empty((n, hwpr), dtype=torch.int64).pin_memory()
h_lens = torch.empty((n,), dtype=torch.int16).pin_memory()
h_packed.copy_(d_packed[:, :hwpr]); h_lens.copy_(d_lens)
torch.cuda.synchronize()
stream = torch.cuda.current_stream()
g = gpu.GpuBuildGraph(0)
g.set_stream(stream.cuda_stream)
h_edges = h_crows = None


ASYNC = False


def step(sync_each):
    global h_edges, h_crows
    t = [time.perf_counter()]

    def mark():
        if sync_each:
            torch.cuda.synchronize()
        t.append(time.perf_counter())
    if ASYNC:
        g.load_reads_async(h_packed.data_ptr(), h_lens.data_ptr(), n, hwpr, 150, 150)
    else:
        g.load_reads_ptr(h_packed.data_ptr(), h_lens.data_ptr(), n, hwpr)
    mark()
    g.build_graph(50, 4); mark()
    nc, ne = g.counts()
    if h_edges is None:
        h_edges = torch.empty((int(ne * 1.1) + 16, 4), dtype=torch.int32).pin_memory()
        g.set_edge_sink(h_edges.data_ptr(), h_edges.shape[0])
        h_crows = torch.empty((int(nc * 1.1) + 16, 4), dtype=torch.int32).pin_memory()
    e = g.edges(out=h_edges.numpy().view(gpu.EDGE_DTYPE).reshape(-1)); mark()
    c = g.contained_into(h_crows.numpy().view(gpu.CROW_DTYPE).reshape(-1)); mark()
    torch.cuda.synchronize()
    t.append(time.perf_counter())
    return [1000 * (b - a) for a, b in zip(t, t[1:])], g.stats()["ms_total"]


for ASYNC in (False, True):
  for _ in range(3):
    step(False)
  for sync_each in (True, False):
    acc = None
    for _ in range(5):
        ms, dev_ms = step(sync_each)
        acc = ms if acc is None else [a + b for a, b in zip(acc, ms)]
    print("load_reads_async" if ASYNC else "load_reads", "|", "sync after each stage" if sync_each else "as bench.py runs it", "| load_reads %.2f  build_graph %.2f  edges %.2f  contained %.2f  tail %.2f  | sum %.2f ms, device path %.2f ms"
          % (*[a / 5 for a in acc], sum(acc) / 5, dev_ms), flush=True)
# the host loop over the lengths inside disco_gpu_load_reads (min / max)
import numpy as np
a = h_lens.numpy()
t0 = time.perf_counter(); mn, mx = int(a.min()), int(a.max()); t1 = time.perf_counter()
print("numpy min+max over %d lengths: %.2f ms" % (n, 1000 * (t1 - t0)))
t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
tmp = torch.empty((n, hwpr), dtype=torch.int64, device=dev)
t0.record(); tmp.copy_(h_packed, non_blocking=True); t1.record(); torch.cuda.synchronize()
print("H2D of %.0f MB pinned: %.2f ms = %.1f GB/s" % (h_packed.numel() * 8 / 1e6, t0.elapsed_time(t1), h_packed.numel() * 8 / 1e6 / t0.elapsed_time(t1)))
