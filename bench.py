#!/usr/bin/env python
"""bench.py -- BuildGraph hot path on B200: reads/s overlap-searched (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--reads R] [--impl ours|reference]

One step = one pass of the whole hot path (hash table -> contained reads -> hash table of the survivors -> overlap
search -> transitive reduction) over one batch of synthetic reads.  `value` times the device path with the packed reads
already resident in HBM; `e2e` times the C-ABI call with HOST buffers (pinned): H2D of the packed reads, the same device
path, D2H of the contained rows and the reduced edge list.

Workloads (BASELINE.json configs): the headline line is config 2's shape at every N (150 bp, 30x, one random genome,
minOverlap 50, 10 M reads per GPU: weak scaling); the same JSON line carries, under "config3_shape", a second measurement on
config 3's shape (200 genomes with log-normal abundances, 12.5 M reads per GPU = 100 M reads on 8 GPUs).

N > 1: reads sharded by read id, table and reads replicated on every GPU (the BuildGraphMPI partitioning) or, with
--partition key-sharded, the table sharded by key (BuildGraphMPIRMA's).  Every multi-GPU measurement is followed, outside
the timed region, by a parity check: rank 0 runs the single-GPU path on the whole read set and the order-independent
checksums of the two results (edges: count / sum / xor of a 64-bit mix of (src, dst, orientation, offset); contained rows
the same) must agree, else the process exits non-zero.  At N = 1 the parity record is the comparison of the GPU result
with the files the reference binary wrote for the cpu_baseline sample.
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
MASK64 = (1 << 64) - 1


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
    n_genomes = {"metagenome": 200, "metagenome2000": 2000}.get(workload, 1)   # config 3 / config 5 shape: log-normal abundances, sigma = 1
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
# order-independent checksums of a result (parity_check)
def _mix64(x):
    x = x.copy()
    x ^= x >> np.uint64(33); x *= np.uint64(0xff51afd7ed558ccd)
    x ^= x >> np.uint64(33); x *= np.uint64(0xc4ceb9fe1a85ec53)
    x ^= x >> np.uint64(33)
    return x


def checksum4(a, b, c, d):
    """(count, sum mod 2^64, xor) of a 64-bit mix of four u32 columns"""
    if len(a) == 0:
        return (0, 0, 0)
    with np.errstate(over="ignore"):
        k = _mix64((a.astype(np.uint64) << np.uint64(32)) | b.astype(np.uint64))
        k = _mix64(k + ((c.astype(np.uint64) << np.uint64(32)) | d.astype(np.uint64)) * np.uint64(0x9E3779B97F4A7C15))
        return (int(len(k)), int(np.add.reduce(k, dtype=np.uint64)), int(np.bitwise_xor.reduce(k)))


def edge_checksum(e):
    return checksum4(e["src"], e["dst"], e["orient"], e["offset"])


def crow_checksum(c):
    return checksum4(c["contained"], c["container"], c["orient"], c["start"])


def combine(sums):
    n = s = x = 0
    for a, b, c in sums:
        n += a; s = (s + b) & MASK64; x ^= c
    return (n, s, x)


# ---------------------------------------------------------------------------------------------------------------------
def _ref_sample(sample_reads):
    """Runs the reference binary on `sample_reads` reads of the headline generator; returns (result dict, fasta path)."""
    from oracle import oracle
    if not oracle.have_ref():
        try:
            oracle.build()
        except Exception:
            pass
    if not oracle.have_ref():
        return None, None, None
    cores = os.cpu_count() or 1
    rs = make_reads(sample_reads, seed=2)
    d = tempfile.mkdtemp(prefix="disco_ref_")
    fa = os.path.join(d, "reads.fa")
    rs.write_fasta(fa)
    return oracle, fa, cores


def run_reference(args):
    """The reference's own OpenMP BuildGraph (oracle/_ref/buildG = unmodified algorithm + the two SURVEY 8c patches),
    all host cores, on a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = args.ref_reads
    line = {"impl": "reference", "metric": METRIC, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
            "data": "synthetic",
            "config": {"workload": f"synthetic {sample} x {READ_LEN}bp single-genome reads, {COVERAGE:.0f}x, minOverlap={MIN_OVERLAP} "
                                   f"(bounded sample of config 2: 10M x 150bp; the one full-size CPU run is recorded in "
                                   f"profiles/r02_ref_full_config2.json)", "min_overlap": MIN_OVERLAP}}
    oracle, fa, cores = _ref_sample(sample)
    if oracle is None:
        line["unavailable"] = "oracle/_ref/buildG missing (reference not built in this snapshot)"
        print(json.dumps(line), flush=True)
        return
    d = os.path.dirname(fa)
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


def cpu_baseline(sample_reads, device):
    """Bounded CPU sample timed beside the GPU run (rank 0, N=1) -- and compared with it: the GPU path runs the same FASTA
    through the same front end (parser + filter), and its canonical edge lines and contained rows must equal the files the
    reference wrote.  Returns (cpu_baseline dict, parity dict)."""
    cores = os.cpu_count() or 1
    try:
        oracle, fa, cores = _ref_sample(sample_reads)
        if oracle is None:
            return None, None
        d = os.path.dirname(fa)
        r = oracle.run_ref([fa], os.path.join(d, "o"), MIN_OVERLAP, threads=cores, mem_gb=64)
        t = r["times"]["buildOverlapGraphFromHashTable"] + r["times"].get("insertDataset", 0.0)
        base = {"value": sample_reads / t, "unit": UNIT, "cores": cores, "kind": "reference",
                "sample": f"{sample_reads} reads of the same generator; oracle/_ref/buildG -t {cores}: insertDataset "
                          f"{r['times'].get('insertDataset', 0.0):.2f}s + buildOverlapGraphFromHashTable {r['times']['buildOverlapGraphFromHashTable']:.2f}s"}
    except Exception as e:  # never let the baseline leg kill the bench line
        return {"value": None, "unit": UNIT, "cores": cores, "kind": "reference", "sample": f"failed: {e}"}, None
    try:
        from disco_b200.buildgraph import BuildGraph
        bg = BuildGraph(min_overlap=MIN_OVERLAP, device=device)
        bg.add_file(fa)
        res = bg.run()
        mine = sorted(bg.edge_lines())
        rows = sorted(bg.crow_lines())
        st = res.stats
        bg.close()
        ref_rows = sorted(r["contained_rows"])
        contained_ok = set(int(x.split("\t")[0]) for x in rows) == r["contained_set"]
        exact_claim = st["cap_fired"] == 0 and st["multi_overlap_pairs"] == 0 and st["one_sided_edges"] == 0
        parity = {"against": f"oracle/_ref/buildG -t {cores} on the cpu_baseline sample ({sample_reads} reads)",
                  "edges": len(mine), "edges_ref": len(r["edges"]), "edges_equal": mine == r["edges"],
                  "contained": len(rows), "contained_ref": len(ref_rows), "contained_set_equal": contained_ok,
                  "contained_rows_equal_as_sets": rows == ref_rows,   # row attribution is pinned by -t 1 only
                  "cap_fired": int(st["cap_fired"]), "exactness_claimed": bool(exact_claim)}
        parity["ok"] = bool(parity["edges_equal"] and contained_ok) if exact_claim else bool(contained_ok)
    except Exception as e:
        parity = {"ok": False, "error": str(e)}
    return base, parity


# ---------------------------------------------------------------------------------------------------------------------
def measure(args, workload, reads_per_gpu, world, rank, local, primary):
    """One workload: device-resident timing, end-to-end timing, roofline, parity.  Returns the JSON object (rank 0)."""
    import torch
    import torch.distributed as dist
    from disco_b200 import gpu, multigpu

    dev = torch.device("cuda", local)
    n = reads_per_gpu * world             # weak scaling: per-GPU query share is fixed
    wpr = 8                               # 64-byte rows
    d_packed, d_lens = make_packed_on_gpu(n, {"single": 2, "metagenome": 3, "metagenome2000": 5}[workload], dev, wpr, workload)
    # pinned host copies (the reference-facing call takes host memory): compact rows, 5 words per 150-bp read -- the
    # library re-strides on the device, so PCIe carries 40 instead of 64 bytes per read
    hwpr = (READ_LEN + 31) // 32
    lo, hi = (rank * n) // world, ((rank + 1) * n) // world
    if world > 1:   # each rank uploads its own shard only
        h_packed = torch.empty((hi - lo, hwpr), dtype=torch.int64).pin_memory()
        h_lens = torch.empty((hi - lo,), dtype=torch.int16).pin_memory()
        h_packed.copy_(d_packed[lo:hi, :hwpr]); h_lens.copy_(d_lens[lo:hi])
    else:
        h_packed = torch.empty((n, hwpr), dtype=torch.int64).pin_memory()
        h_lens = torch.empty((n,), dtype=torch.int16).pin_memory()
        h_packed.copy_(d_packed[:, :hwpr]); h_lens.copy_(d_lens)
    torch.cuda.synchronize()
    stream = torch.cuda.current_stream()
    g = gpu.GpuBuildGraph(local)
    g.set_stream(stream.cuda_stream)
    key_sharded = args.partition == "key-sharded"
    runner = None
    if world > 1:
        if args.partition == "replicated-gather":
            runner = multigpu.ShardedBuildGraph(g, rank, world)
        else:   # adjacency partitioned by query range, read through peer pointers; table sharded by key or replicated
            runner = multigpu.KeyShardedBuildGraph(g, rank, world, shard_table=key_sharded)

    def device_step():
        # the packed reads are resident in HBM in the library's row layout: used in place, no copy
        g.use_reads_device(d_packed.data_ptr(), d_lens.data_ptr(), n, wpr, READ_LEN, READ_LEN)
        if runner:
            runner.build_graph(MIN_OVERLAP, 4)
        else:
            g.build_graph(MIN_OVERLAP, 4)

    h_edges = None
    h_crows = None
    d_in_packed = torch.empty((n, hwpr), dtype=torch.int64, device=dev) if world > 1 else None
    d_in_lens = torch.empty_like(d_lens) if world > 1 else None

    trace_e2e = bool(os.environ.get("DISCO_TRACE_E2E"))
    marks = []

    def mark(name):
        if trace_e2e:
            e = torch.cuda.Event(enable_timing=True)
            e.record(stream)
            marks.append((name, e, time.perf_counter()))

    def e2e_step():
        nonlocal h_edges, h_crows
        marks.clear()
        mark("start")
        if world > 1:
            # every rank uploads only its own shard of the packed reads over PCIe and the shards are all-gathered over
            # NVLink (the read set is replicated on every GPU in this partitioning)
            d_in_packed[lo:hi].copy_(h_packed, non_blocking=True)
            d_in_lens[lo:hi].copy_(h_lens, non_blocking=True)
            mark("h2d")
            dist.all_gather_into_tensor(d_in_packed.view(-1), d_in_packed[lo:hi].view(-1))
            lb = d_in_lens.view(torch.uint8)   # NCCL has no int16
            dist.all_gather_into_tensor(lb, lb[2 * lo:2 * hi])
            mark("all-gather")
            g.load_reads_device(d_in_packed.data_ptr(), d_in_lens.data_ptr(), n, hwpr, READ_LEN, READ_LEN)
            mark("restride")
        else:
            # pinned host rows; the upload runs inside build_graph, chunk by chunk under the table build (the loader
            # knows the shortest / longest read, as the reference's Dataset does)
            g.load_reads_async(h_packed.data_ptr(), h_lens.data_ptr(), n, hwpr, READ_LEN, READ_LEN)
        if runner:
            runner.build_graph(MIN_OVERLAP, 4)
        else:
            g.build_graph(MIN_OVERLAP, 4)
        mark("build_graph")
        nc, ne = g.counts()
        if h_edges is None or h_edges.shape[0] < ne:
            h_edges = torch.empty((int(ne * 1.1) + 16, 4), dtype=torch.int32).pin_memory()
            g.set_edge_sink(h_edges.data_ptr(), h_edges.shape[0])   # from the next call on the emission kernel fills it itself
        if h_crows is None or h_crows.shape[0] < nc:
            h_crows = torch.empty((int(nc * 1.1) + 16, 4), dtype=torch.int32).pin_memory()
        e = g.edges(out=h_edges.numpy().view(gpu.EDGE_DTYPE).reshape(-1))
        mark("edges")
        if world > 1:
            # every rank holds all contained rows; like the edges, each rank hands out those of its own read range (the
            # range the driver cut for the edge pass), so the rows cross PCIe once, not once per GPU
            b = getattr(runner, "bounds", None)
            blo, bhi = (b[rank], b[rank + 1]) if b else (lo, hi)
            c = g.contained_range_into(h_crows.numpy().view(gpu.CROW_DTYPE).reshape(-1), blo, bhi)
        else:
            c = g.contained_into(h_crows.numpy().view(gpu.CROW_DTYPE).reshape(-1))
        mark("contained rows")
        if trace_e2e:
            torch.cuda.synchronize()
            sys.stderr.write(f"e2e rank {rank}: " + ", ".join(f"{b[0]} {a[1].elapsed_time(b[1]):.2f} ms (host {1000 * (b[2] - a[2]):.2f})" for a, b in zip(marks, marks[1:]))
                             + f" | phases {({k: round(v, 2) for k, v in g.stats().items() if k.startswith('ms_') and not k.endswith('kernel') and 'edges_' not in k})}\n")
        return e, c

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

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
    ms_step = max_over_ranks(ms_total) / args.steps
    value = n / (ms_step / 1000.0)

    # end to end through host buffers
    for _ in range(2):
        e2e_step()
    barrier()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record(stream)
    for _ in range(args.steps):
        e_out, c_out = e2e_step()
    t1.record(stream)
    barrier()
    ms_e2e = max_over_ranks(t0.elapsed_time(t1)) / args.steps
    h2d = h_packed.numel() * 8 + h_lens.numel() * 2   # per rank: its shard (N > 1) or everything (N = 1)
    d2h = (len(e_out) + len(c_out)) * 16

    st = stats_acc[-1]
    tot_raw, tot_edges = st["raw_directed_edges"], st["n_edges"]
    if world > 1:
        t = torch.tensor([tot_raw, tot_edges], device="cuda", dtype=torch.int64)
        dist.all_reduce(t)
        tot_raw, tot_edges = int(t[0]), int(t[1])

    # ---- parity (outside every timed region) --------------------------------------------------------------------------
    parity = None
    if world > 1 and args.no_parity:
        del runner
        g.close()
        good_everywhere = True
        parity = {"ok": None, "skipped": "--no-parity (the whole read set does not fit one GPU's single-table run at this size)"}
    elif world > 1:
        mine = (edge_checksum(e_out), crow_checksum(c_out))
        every = [None] * world
        dist.all_gather_object(every, mine)
        multi_edges = combine([m[0] for m in every])
        multi_crows = combine([m[1] for m in every])                # each rank handed out the rows of its own read range
        crows_agree = True
        ok = torch.zeros(1, dtype=torch.int64, device=dev)
        # free this rank's multi-GPU buffers before rank 0 takes the whole read set on its own
        del runner
        g.close()
        h_edges = h_crows = d_in_packed = d_in_lens = None
        torch.cuda.empty_cache()
        if rank == 0:
            g1 = gpu.GpuBuildGraph(local)
            g1.set_stream(stream.cuda_stream)
            g1.load_reads_device(d_packed.data_ptr(), d_lens.data_ptr(), n, wpr, READ_LEN, READ_LEN)
            g1.build_graph(MIN_OVERLAP, 4)
            s1 = g1.stats()
            one_edges, one_crows = edge_checksum(g1.edges()), crow_checksum(g1.contained())
            g1.close()
            good = one_edges == multi_edges and one_crows == multi_crows and crows_agree
            parity = {"against": f"single-GPU path on rank 0 over all {n} reads (order-independent checksums: count, sum, xor of a 64-bit mix)",
                      "ok": bool(good), "edges": multi_edges[0], "edges_single_gpu": one_edges[0],
                      "edge_checksum": [hex(multi_edges[1]), hex(multi_edges[2])],
                      "edge_checksum_single_gpu": [hex(one_edges[1]), hex(one_edges[2])],
                      "contained": multi_crows[0], "contained_single_gpu": one_crows[0],
                      "contained_checksum_equal": bool(one_crows == multi_crows), "contained_rows": "each rank the rows of its read range; union compared",
                      "single_gpu_ms_same_input": float(s1["ms_total"]),
                      "single_gpu_reads_per_s_same_input": n / (s1["ms_total"] / 1000.0)}
            ok[0] = 1 if good else 0
        dist.broadcast(ok, 0)
        good_everywhere = bool(int(ok[0]))
    else:
        # next row of the scope table, outside the timed region: parsimplify's contraction + dead-end removal on
        # the edges still in HBM (parity: tests/test_simplify_gpu.py against the reference parsimplify)
        simplify = None
        try:
            ce, inner, sst = g.simplify(MIN_OVERLAP)
            simplify = {"ms": float(sst["ms"]), "rounds": int(sst["rounds"]), "composite_edges": int(len(ce)), "inner_reads": int(len(inner)),
                        "reduced_edges_in": int(st["n_edges"]), "removed_with_dead_ends": int(sst["removed_edges"]),
                        "longest_edge_bases": int((ce["offset_total"]).max()) + READ_LEN if len(ce) else 0}
        except Exception as e:
            simplify = {"error": str(e)}
        g.close()
        good_everywhere = True

    peak, peak_src = measured_peak()
    # Rooflines.  Algorithmic bytes follow SURVEY 8(d), taken from the kernels' own counters of THIS rank's launch
    # (R = 40 B packed read, one filter word per probe, S = 32 B bucket sector, 64-byte candidate rows, 8 B per candidate /
    # adjacency entry); time = CUDA events recorded on the launching stream inside the C ABI (disco_stats.ms_*).  The
    # kernel that takes longest is reported as `roofline`, the others next to it.
    def mean_ms(k):
        return float(np.mean([x[k] for x in stats_acc]))
    ms_probe, ms_verify, ms_pass = mean_ms("ms_edges_probe"), mean_ms("ms_edges_verify"), mean_ms("ms_edges_kernel")
    ms_mark, ms_emit = mean_ms("ms_mark_kernel"), mean_ms("ms_emit_kernel")
    q, pr, bk, vf, en = st["queries_edges"], st["probes_edges"], st["buckets_edges"], st["verified_edges"], st["raw_directed_edges"]
    alg = {
        "k_probe_flat": q * 40 + pr * 4 + bk * 32 + vf * 8,             # read, filter word per probe, buckets, candidates written
        "k_verify_flat": q * 40 + vf * 8 + vf * 64 + en * 8,            # read, candidates read, candidate rows, adjacency entries
        "k_reduce_mark": en * 8 + st["mark_entries_fetched"] * 8 + en * 8,   # own rows, visited neighbours' rows, marks written back
        "k_reduce_emit": en * 8 + (st["reduce_entries_fetched"] - st["mark_entries_fetched"]) * 8 + st["n_edges"] * 16,
    }
    alg_pass = q * 40 + pr * 32 + vf * 64 + en * 8                      # SURVEY 8(d) verbatim: R + P*S + H*2S + 8*E_raw
    ms_of = {"k_probe_flat": ms_probe, "k_verify_flat": ms_verify, "k_reduce_mark": ms_mark, "k_reduce_emit": ms_emit}
    traffic, traffic_src = {}, None
    tp = os.path.join(ROOT, "profiles", "traffic_r02.json")
    if os.path.exists(tp) and primary and world == 1:
        try:
            tj = json.load(open(tp))
            if int(tj.get("reads", -1)) == n:
                traffic = tj["dram_bytes_per_launch"]  # dram__bytes_read.sum + dram__bytes_write.sum per kernel, ncu
                traffic_src = tj.get("source")
        except Exception:
            pass

    def roof(name, a, ms):
        ach = a / (ms / 1000.0) / 1e9 if ms > 0 else None
        return {"bound": "hbm", "kernel": name, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": (ach / peak) if ach else None,
                "traffic": traffic.get(name), "traffic_source": traffic_src if traffic.get(name) else None,
                "peak_source": peak_src, "ms_per_launch": ms, "algorithmic_bytes_per_launch": int(a)}
    longest = max(ms_of, key=lambda k: ms_of[k])
    wl_text = {"single": "single-genome", "metagenome": "200-genome log-normal metagenome (config 3 shape)",
               "metagenome2000": "2000-genome log-normal metagenome (config 5 shape)"}[workload]
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
        "data": "synthetic",
        "config": {"workload": f"synthetic {n} x {READ_LEN}bp {wl_text} reads ({COVERAGE:.0f}x mean, both strands, error-free), "
                               f"minOverlap={MIN_OVERLAP}" + (f", {world} GPUs: queries sharded by read id, " + ("reads replicated, table sharded by key and adjacency by query range (remote shards read over NVLink)" if key_sharded else ("table+reads replicated, adjacency all-gathered" if args.partition == "replicated-gather" else "table+reads replicated, adjacency partitioned by query range (neighbours' rows read over NVLink)")) if world > 1 else (" (BASELINE config 2)" if workload == "single" and n == 10_000_000 else "")),
                   "partition": (args.partition if world > 1 else "single"),
                   "reads": n, "reads_per_gpu": reads_per_gpu, "read_len": READ_LEN, "min_overlap": MIN_OVERLAP,
                   "max_edge_per_kmer": 4, "l2": "inputs larger than L2 (packed reads + table > 126 MB), no flush needed"},
        "e2e": {"value": n / (ms_e2e / 1000.0), "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": ms_e2e},
        "gpu_launches": int(sum(x["kernel_launches"] for x in stats_acc)),   # counted by the library's launchers, this rank
        "clocks": clk.summary(),
        "roofline": roof(longest, alg[longest], ms_of[longest]),
        "roofline_other": {k: roof(k, alg[k], ms_of[k]) for k in ms_of if k != longest},
        "roofline_edge_pass": roof("k_probe_flat+k_verify_flat+k_edges_exact", alg_pass, ms_pass),
        # the yardstick that fits this path: random-access rates measured with profiles/gather_bench.cu on the same B200
        # (profiles/r01_gather_microbench.txt): 39.4 G/s for 32-byte accesses (buckets), 22.3 G/s for 64-byte ones (read rows)
        "random_access_peak": {"accesses_per_s_32B": 39.4e9, "accesses_per_s_64B": 22.3e9,
                               "source": "profiles/gather_bench.cu on B200 (profiles/r01_gather_microbench.txt)",
                               "verify_rows_per_s": vf / (ms_verify / 1000.0) if ms_verify > 0 else None,
                               "verify_frac_of_64B_peak": (vf / (ms_verify / 1000.0) / 22.3e9) if ms_verify > 0 else None,
                               "probe_buckets_per_s": bk / (ms_probe / 1000.0) if ms_probe > 0 else None,
                               "probe_frac_of_32B_peak": (bk / (ms_probe / 1000.0) / 39.4e9) if ms_probe > 0 else None},
        "edges_per_s": {"raw_directed": tot_raw / (ms_step / 1000.0), "reduced": tot_edges / (ms_step / 1000.0)},
        "phase_ms": {k: float(np.mean([s[k] for s in stats_acc])) for k in st if k.startswith("ms_")},
        "counters": {k: int(v) for k, v in st.items() if not k.startswith("ms_")},
    }
    if parity is not None:
        line["parity_check"] = parity
    if world == 1 and simplify is not None:
        line["simplify_next_row"] = simplify
    del d_packed, d_lens
    torch.cuda.empty_cache()
    return line, good_everywhere


def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    line, ok = measure(args, args.workload, args.reads, world, rank, local, primary=True)
    if args.config3 and args.workload == "single":
        sub, ok3 = measure(args, "metagenome", args.config3_reads, world, rank, local, primary=False)
        keep = ("value", "ms_per_step", "config", "e2e", "roofline", "phase_ms", "counters", "edges_per_s", "parity_check", "gpu_launches")
        line["config3_shape"] = {k: sub[k] for k in keep if k in sub}
        ok = ok and ok3
    if rank == 0 and world == 1 and not args.no_cpu:
        base, parity = cpu_baseline(args.ref_reads, local)
        line["cpu_baseline"] = base
        if parity is not None:
            line["parity_check"] = parity
            ok = ok and bool(parity.get("ok"))
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    if not ok:
        sys.stderr.write("bench.py: PARITY CHECK FAILED (see parity_check in the JSON line)\n")
        sys.exit(3)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--reads", type=int, default=10_000_000, help="reads per GPU (config 2: 10M)")
    ap.add_argument("--ref-reads", type=int, default=400_000, help="reads in the bounded CPU sample")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (and its parity comparison)")
    ap.add_argument("--no-parity", action="store_true", help="N > 1: skip the single-GPU parity run (sizes whose single-table run does not fit one GPU)")
    ap.add_argument("--workload", default="single", choices=["single", "metagenome", "metagenome2000"],
                    help="single = BASELINE config 2's shape (headline); metagenome = config 3's shape (200 genomes, log-normal abundance)")
    ap.add_argument("--no-config3", dest="config3", action="store_false", help="skip the second measurement on config 3's shape")
    ap.add_argument("--config3-reads", type=int, default=12_500_000, help="reads per GPU of the config-3-shape measurement (8 GPUs: 100M)")
    ap.add_argument("--partition", default=os.environ.get("DISCO_PARTITION", "replicated"), choices=["replicated", "replicated-gather", "key-sharded"],
                    help="N > 1 only: replicated = table + reads replicated, queries sharded by read id (config 3; the adjacency stays "
                         "with its owner), replicated-gather = the same with the adjacency all-gathered, key-sharded = table sharded by "
                         "key (config 5's partitioning)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
