"""2-GPU run (torchrun, NCCL) of the sharded path against the single-GPU result.  Skipped with fewer than 2 GPUs."""
import json
import os
import subprocess
import sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu

SCRIPT = r'''
import os, sys, json
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, %r)
from disco_b200 import gpu, host, synth, multigpu
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rs = synth.dup_contained(30000, 150, 40.0, seed=77)
packed, lens = host.pack_codes(rs.codes, rs.off)
g = gpu.GpuBuildGraph(local)
g.set_stream(torch.cuda.current_stream().cuda_stream)
g.load_reads(packed, lens)
drv = %s
drv.build_graph(50, 4)
e = g.edges(); c = g.contained()
b = getattr(drv, "bounds", None)           # ranges with equal numbers of non-contained reads, or the plain partition
lo, hi = (b[rank], b[rank + 1]) if b else multigpu.partition(rs.n, rank, world)
assert ((e["src"] >= lo) & (e["src"] < hi)).all()          # each rank emits the edges whose lower endpoint it owns
gathered = [None] * world
dist.all_gather_object(gathered, (e.tobytes(), c.tobytes()))
if rank == 0:
    edges = np.concatenate([np.frombuffer(x[0], dtype=gpu.EDGE_DTYPE) for x in gathered])
    g1 = gpu.GpuBuildGraph(local); g1.load_reads(packed, lens); g1.build_graph(50, 4)
    e1 = gpu.sort_edges(g1.edges()); c1 = g1.contained()
    ok_e = np.array_equal(gpu.sort_edges(edges), e1)
    ok_c = all(np.array_equal(np.sort(np.frombuffer(x[1], dtype=gpu.CROW_DTYPE), order=["contained"]), np.sort(c1, order=["contained"])) for x in gathered)
    print("RESULT", json.dumps({"edges_equal": bool(ok_e), "contained_equal": bool(ok_c), "n_edges": int(len(e1)), "n_contained": int(len(c1))}))
dist.destroy_process_group()
'''


@pytest.mark.parametrize("driver", ["multigpu.ShardedBuildGraph(g, rank, world)", "multigpu.KeyShardedBuildGraph(g, rank, world)",
                                    "multigpu.KeyShardedBuildGraph(g, rank, world, shard_table=False)",
                                    "multigpu.KeyShardedBuildGraph(g, rank, world, shard_table=False, balance=False)"],
                         ids=["modeA_gathered", "modeB_key_sharded", "replicated_table_partitioned_adjacency", "the_same_unbalanced"])
def test_two_gpus_match_one(tmp_path, driver):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "run.py"
    script.write_text(SCRIPT % (ROOT, driver))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29611", str(script)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("RESULT")][0]
    res = json.loads(line[7:])
    assert res["edges_equal"] and res["contained_equal"] and res["n_edges"] > 0 and res["n_contained"] > 0
