#!/usr/bin/env python
"""bench.py -- BuildGraph hot path on B200: reads/s overlap-searched (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--reads R] [--impl ours|reference]

One step = one pass of the whole hot path (hash table -> contained reads -> hash table of the survivors -> overlap
search -> transitive reduction) over one batch of synthetic reads of BASELINE config 2's shape (150 bp, 30x,
single random genome, minOverlap 50).  `value` times the device path with the packed reads already resident in HBM;
`e2e` times the C-ABI call with HOST buffers (pinned): H2D of the packed reads, the same device path, D2H of the
contained rows and the reduced edge list.  N > 1: reads sharded by read id, table and reads replicated on every GPU
(the BuildGraphMPI partitioning); per-GPU work is fixed (weak scaling), no collective on the timed path except the
containment-key all-reduce and the adjacency all-gather the algorithm needs.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "reads/sec overlap-searched"
UNIT = "reads/s"
MIN_OVERLAP = 50
READ_LEN = 150
COVERAGE = 30.0


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index=0):
        self.index = index
        self.rows = []
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for nme, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_reads(n_reads, seed):
    from disco_b200 import synth
    return synth.single_genome(n_reads, READ_LEN, COVERAGE, seed=seed)


def make_packed_on_gpu(n_reads, seed, device, wpr, workload="single"):
    """Same shape as synth.single_genome (uniform-random genome, uniform starts, strand ~ Bernoulli(1/2), error-free),
    generated and 2-bit packed with torch on the device so that the 8-GPU runs (80M reads per rank) start in seconds.
    Data generation is outside every timed region."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    glen = max(READ_LEN + 1, int(round(n_reads * READ_LEN / COVERAGE)))
    genome = torch.randint(0, 4, (glen,), dtype=torch.uint8, device=device, generator=g)
    n_genomes = 200 if workload == "metagenome" else 1       # config 3 shape: log-normal abundances, sigma = 1
    gl = glen // n_genomes
    if n_genomes > 1:
        ab = torch.exp(torch.randn((n_genomes,), device=device, generator=g))
        ab = ab / ab.sum()
    out = torch.zeros((n_reads, wpr), dtype=torch.int64, device=device)
    words = (READ_LEN + 31) // 32
    shifts = (62 - 2 * torch.arange(32, device=device, dtype=torch.int64))
    ar = torch.arange(READ_LEN, device=device, dtype=torch.int64)
    step = 1 << 20
    for lo in range(0, n_reads, step):
        m = min(step, n_reads - lo)
        if n_genomes > 1:
            which = torch.multinomial(ab, m, replacement=True, generator=g)
            starts = which * gl + torch.randint(0, gl - READ_LEN + 1, (m,), device=device, generator=g, dtype=torch.int64)
        else:
            starts = torch.randint(0, glen - READ_LEN + 1, (m,), device=device, generator=g, dtype=torch.int64)
        flip = torch.rand((m,), device=device, generator=g) < 0.5
        codes = genome[starts[:, None] + ar[None, :]]
        codes = torch.where(flip[:, None], 3 - codes.flip(1), codes).to(torch.int64)
        codes = torch.nn.functional.pad(codes, (0, words * 32 - READ_LEN))
        out[lo:lo + m, :words] = (codes.view(m, words, 32) << shifts).sum(dim=2)   # disjoint bit fields: sum == or
    lens = torch.full((n_reads,), READ_LEN, dtype=torch.int16, device=device)
    return out, lens


# ---------------------------------------------------------------------------------------------------------------------
def run_reference(args):
    """The reference's own OpenMP BuildGraph (oracle/_ref/buildG = unmodified algorithm + the two SURVEY 8c patches),
    all host cores, on a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle
    cores = os.cpu_count() or 1
    sample = args.ref_reads
    line = {"impl": "reference", "metric": METRIC, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
            "data": "synthetic",
            "config": {"workload": f"synthetic {sample} x {READ_LEN}bp single-genome reads, {COVERAGE:.0f}x, minOverlap={MIN_OVERLAP} "
                                   f"(bounded sample of config 2: 10M x 150bp)", "min_overlap": MIN_OVERLAP}}
    if not oracle.have_ref():
        try:
            oracle.build()
        except Exception:
            pass
    if not oracle.have_ref():
        line["unavailable"] = "oracle/_ref/buildG missing (reference not built in this snapshot)"
        print(json.dumps(line), flush=True)
        return
    rs = make_reads(sample, seed=2)
    d = tempfile.mkdtemp(prefix="disco_ref_")
    fa = os.path.join(d, "reads.fa")
    rs.write_fasta(fa)
    times = []
    for it in range(args.warmup + args.steps):
        pre = os.path.join(d, f"run{it}", "o")
        r = oracle.run_ref([fa], pre, MIN_OVERLAP, threads=cores, mem_gb=64)
        t = r["times"].get("buildOverlapGraphFromHashTable", None)
        t_ins = r["times"].get("insertDataset", 0.0)
        if t is None:
            line["unavailable"] = "reference run failed: " + r["log"][-200:].replace("\n", " ")
            print(json.dumps(line), flush=True)
            return
        if it >= args.warmup:
            times.append(t + t_ins)  # hash table build + graph stage = the same span our step covers
    ms = 1000.0 * float(np.mean(times))
    v = sample / (ms / 1000.0)
    line.update({"value": v, "ms_per_step": ms,
                 "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "reference",
                                  "sample": f"{sample} reads of the same generator (seed 2); insertDataset + buildOverlapGraphFromHashTable wall time, -t {cores}"},
                 "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    print(json.dumps(line), flush=True)


def cpu_baseline(sample_reads):
    """Bounded CPU sample timed beside the GPU run (rank 0, N=1)."""
    from oracle import oracle
    cores = os.cpu_count() or 1
    try:
        if not oracle.have_ref():
            oracle.build()
        if not oracle.have_ref():
            return None
        rs = make_reads(sample_reads, seed=2)
        d = tempfile.mkdtemp(prefix="disco_cpu_")
        fa = os.path.join(d, "reads.fa")
        rs.write_fasta(fa)
        r = oracle.run_ref([fa], os.path.join(d, "o"), MIN_OVERLAP, threads=cores, mem_gb=64)
        t = r["times"]["buildOverlapGraphFromHashTable"] + r["times"].get("insertDataset", 0.0)
        return {"value": sample_reads / t, "unit": UNIT, "cores": cores, "kind": "reference",
                "sample": f"{sample_reads} reads of the same generator; oracle/_ref/buildG -t {cores}: insertDataset "
                          f"{r['times'].get('insertDataset', 0.0):.2f}s + buildOverlapGraphFromHashTable {r['times']['buildOverlapGraphFromHashTable']:.2f}s"}
    except Exception as e:  # never let the baseline leg kill the bench line
        return {"value": None, "unit": UNIT, "cores": cores, "kind": "reference", "sample": f"failed: {e}"}


# ---------------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from disco_b200 import gpu, host, multigpu

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n_total = args.reads * world          # weak scaling: per-GPU query share is fixed
    n = n_total
    wpr = 8                               # 64-byte rows
    d_packed, d_lens = make_packed_on_gpu(n, 2, torch.device("cuda", local), wpr, args.workload)
    # pinned host copies (the reference-facing call takes host memory): compact rows, 5 words per 150-bp read -- the
    # library re-strides on the device, so PCIe carries 40 instead of 64 bytes per read
    hwpr = (READ_LEN + 31) // 32
    h_packed = torch.empty((n, hwpr), dtype=torch.int64).pin_memory()
    h_lens = torch.empty((n,), dtype=torch.int16).pin_memory()
    h_packed.copy_(d_packed[:, :hwpr])
    h_lens.copy_(d_lens)
    torch.cuda.synchronize()
    stream = torch.cuda.current_stream()
    g = gpu.GpuBuildGraph(local)
    g.set_stream(stream.cuda_stream)
    # N > 1: Mode A (reads + table replicated, adjacency all-gathered; BASELINE config 3) or Mode B (table sharded by key,
    # adjacency by query range, remote shards read through NVLink; config 5's partitioning)
    key_sharded = args.partition == "key-sharded"
    runner = None
    if world > 1:
        runner = multigpu.KeyShardedBuildGraph(g, rank, world) if key_sharded else multigpu.ShardedBuildGraph(g, rank, world)
    lo, hi = (rank * n) // world, ((rank + 1) * n) // world

    def device_step():
        g.load_reads_device(d_packed.data_ptr(), d_lens.data_ptr(), n, wpr, READ_LEN, READ_LEN)
        if runner:
            runner.build_graph(MIN_OVERLAP, 4)
        else:
            g.build_graph(MIN_OVERLAP, 4)

    h_edges = None
    h_crows = None

    d_in_packed = torch.empty((n, hwpr), dtype=torch.int64, device=d_packed.device) if world > 1 else None
    d_in_lens = torch.empty_like(d_lens) if world > 1 else None

    def e2e_step():
        nonlocal h_edges, h_crows
        if world > 1:
            # every rank uploads only its own shard of the packed reads over PCIe and the shards are all-gathered over
            # NVLink (the read set is replicated on every GPU in this partitioning)
            d_in_packed[lo:hi].copy_(h_packed[lo:hi], non_blocking=True)
            d_in_lens[lo:hi].copy_(h_lens[lo:hi], non_blocking=True)
            dist.all_gather_into_tensor(d_in_packed.view(-1), d_in_packed[lo:hi].view(-1))
            lb = d_in_lens.view(torch.uint8)   # NCCL has no int16
            dist.all_gather_into_tensor(lb, lb[2 * lo:2 * hi])
            g.load_reads_device(d_in_packed.data_ptr(), d_in_lens.data_ptr(), n, hwpr, READ_LEN, READ_LEN)
        else:
            g.load_reads_ptr(h_packed.data_ptr(), h_lens.data_ptr(), n, hwpr)
        if runner:
            runner.build_graph(MIN_OVERLAP, 4)
        else:
            g.build_graph(MIN_OVERLAP, 4)
        nc, ne = g.counts()
        if h_edges is None or h_edges.shape[0] < ne:
            h_edges = torch.empty((int(ne * 1.1) + 16, 4), dtype=torch.int32).pin_memory()
        if h_crows is None or h_crows.shape[0] < nc:
            h_crows = torch.empty((int(nc * 1.1) + 16, 4), dtype=torch.int32).pin_memory()
        e = g.edges(out=h_edges.numpy().view(gpu.EDGE_DTYPE).reshape(-1))
        c = g.contained_into(h_crows.numpy().view(gpu.CROW_DTYPE).reshape(-1))
        return len(e), len(c)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record(stream)
        out = None
        for _ in range(steps):
            out = fn()
        t1.record(stream)
        barrier()
        ms = t0.elapsed_time(t1)
        if world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, out

    for _ in range(max(args.warmup, 3)):
        device_step()
    stats_acc = []
    with ClockSampler(local) as clk:
        # per-step loop so that the per-kernel counters of every step can be read (between steps, outside the kernels)
        barrier()
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record(stream)
        for _ in range(args.steps):
            device_step()
            stats_acc.append(g.stats())
        t1.record(stream)
        barrier()
        ms_total = t0.elapsed_time(t1)
    if world > 1:
        t = torch.tensor([ms_total], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = n / (ms_step / 1000.0)

    # end to end through host buffers
    for _ in range(2):
        e2e_step()
    ms_e2e, (ne_out, nc_out) = timed(e2e_step, args.steps)
    ms_e2e /= args.steps
    h2d = (h_packed.numel() * 8 + h_lens.numel() * 2) // world   # per rank: its shard (N > 1) or everything (N = 1)
    d2h = (ne_out + nc_out) * 16

    st = stats_acc[-1]
    tot_raw, tot_edges = st["raw_directed_edges"], st["n_edges"]
    if world > 1:
        t = torch.tensor([tot_raw, tot_edges], device="cuda", dtype=torch.int64)
        dist.all_reduce(t)
        tot_raw, tot_edges = int(t[0]), int(t[1])
    peak, peak_src = measured_peak()
    # Roofline of the dominant kernel.  The edge pass runs as k_edges_probe (hash, filter, bucket probe) and
    # k_edges_verify (candidate fetch + overlap compare); both take about the same time, verify is the larger one and
    # is reported as `roofline`, probe and the whole pass next to it.  Algorithmic bytes follow SURVEY 8(d)
    # (R = 40 B packed read, S = 32 B sector, candidate rows of 48 B = 2 sectors, 8 B per parked candidate / entry),
    # taken from the kernels' own counters of THIS rank's launch; time = CUDA events recorded on the launching stream
    # inside the C ABI (disco_stats.ms_edges_*).
    def mean_ms(k):
        return float(np.mean([x[k] for x in stats_acc]))
    ms_probe, ms_verify, ms_pass = mean_ms("ms_edges_probe"), mean_ms("ms_edges_verify"), mean_ms("ms_edges_kernel")
    q, pr, bk, vf, en = st["queries_edges"], st["probes_edges"], st["buckets_edges"], st["verified_edges"], st["raw_directed_edges"]
    alg_probe = q * 40 + pr * 4 + bk * 32 + vf * 8                 # read, 1 filter word per probe, buckets, parked candidates
    alg_verify = q * 40 + vf * 8 + vf * 64 + en * 8                # read, parked candidates, candidate rows, adjacency entries
    alg_pass = q * 40 + pr * 32 + vf * 64 + en * 8                 # R + P*S + H*2S + 8*E_raw
    peak, peak_src = measured_peak()

    def roof(name, alg, ms, traffic=None):
        ach = alg / (ms / 1000.0) / 1e9 if ms > 0 else None
        return {"bound": "hbm", "kernel": name, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": (ach / peak) if ach else None,
                "traffic": traffic, "peak_source": peak_src, "ms_per_launch": ms, "algorithmic_bytes_per_launch": int(alg)}
    traffic = {}
    tp = os.path.join(ROOT, "profiles", "traffic_search_edges.json")
    if os.path.exists(tp):
        try:
            tj = json.load(open(tp))
            if int(tj.get("reads", -1)) == n and world == 1:
                traffic = tj["dram_bytes_per_launch"]  # dram__bytes_read.sum + dram__bytes_write.sum per kernel, ncu
        except Exception:
            pass
    if ms_verify <= 0:   # fused single-kernel variant (DISCO_FUSED)
        ms_verify, alg_verify = ms_pass, alg_pass
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
        "data": "synthetic",
        "config": {"workload": f"synthetic {n} x {READ_LEN}bp {'single-genome' if args.workload == 'single' else '200-genome log-normal metagenome'} reads ({COVERAGE:.0f}x mean, both strands, error-free), "
                               f"minOverlap={MIN_OVERLAP}" + (f", {world} GPUs: queries sharded by read id, " + ("reads replicated, table sharded by key and adjacency by query range (remote shards read over NVLink)" if key_sharded else "table+reads replicated") if world > 1 else " (BASELINE config 2 when --reads 10000000)"),
                   "partition": (args.partition if world > 1 else "single"),
                   "reads": n, "reads_per_gpu": args.reads, "read_len": READ_LEN, "min_overlap": MIN_OVERLAP,
                   "max_edge_per_kmer": 4, "l2": "inputs larger than L2 (packed reads + table > 126 MB), no flush needed"},
        "e2e": {"value": n / (ms_e2e / 1000.0), "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": ms_e2e},
        "gpu_launches": int(10 * args.steps),
        "clocks": clk.summary(),
        "roofline": roof("k_edges_verify", alg_verify, ms_verify, traffic.get("k_edges_verify") if isinstance(traffic, dict) else None),
        "roofline_probe": roof("k_edges_probe", alg_probe, ms_probe, traffic.get("k_edges_probe") if isinstance(traffic, dict) else None),
        "roofline_edge_pass": roof("k_edges_probe+k_edges_verify+k_edges_exact", alg_pass, ms_pass,
                                   traffic.get("edge_pass") if isinstance(traffic, dict) else None),
        # the yardstick that fits this path: random-access rates measured with profiles/gather_bench.cu on the same B200
        # (profiles/r01_gather_microbench.txt): 39.4 G/s for 32-byte accesses (buckets), 22.3 G/s for 64-byte ones (read rows)
        "random_access_peak": {"accesses_per_s_32B": 39.4e9, "accesses_per_s_64B": 22.3e9,
                               "source": "profiles/gather_bench.cu on B200 (profiles/r01_gather_microbench.txt)",
                               "verify_rows_per_s": vf / (ms_verify / 1000.0) if ms_verify > 0 else None,
                               "verify_frac_of_64B_peak": (vf / (ms_verify / 1000.0) / 22.3e9) if ms_verify > 0 else None,
                               "probe_buckets_per_s": bk / (ms_probe / 1000.0) if ms_probe > 0 else None,
                               "probe_frac_of_32B_peak": (bk / (ms_probe / 1000.0) / 39.4e9) if ms_probe > 0 else None,
                               "edge_pass_accesses_per_s": (bk + vf) / (ms_pass / 1000.0) if ms_pass > 0 else None},
        "edges_per_s": {"raw_directed": tot_raw / (ms_step / 1000.0), "reduced": tot_edges / (ms_step / 1000.0)},
        "phase_ms": {k: float(np.mean([s[k] for s in stats_acc])) for k in st if k.startswith("ms_")},
        "counters": {k: int(v) for k, v in st.items() if not k.startswith("ms_")},
    }
    if rank == 0 and world == 1 and not args.no_cpu:
        line["cpu_baseline"] = cpu_baseline(args.ref_reads)
    if rank == 0:
        print(json.dumps(line), flush=True)
    g.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--reads", type=int, default=10_000_000, help="reads per GPU (config 2: 10M)")
    ap.add_argument("--ref-reads", type=int, default=400_000, help="reads in the bounded CPU sample")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--workload", default="single", choices=["single", "metagenome"],
                    help="single = BASELINE config 2 (headline); metagenome = config 3 shape (200 genomes, log-normal abundance)")
    ap.add_argument("--partition", default=os.environ.get("DISCO_PARTITION", "replicated"), choices=["replicated", "key-sharded"],
                    help="N > 1 only: replicated = Mode A (config 3), key-sharded = Mode B (config 5's partitioning)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
