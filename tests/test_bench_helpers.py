"""bench.py's parity bookkeeping on CPU: the order-independent checksums every multi-GPU line is judged by, and the
reference arm's contract when the reference binary is missing."""
import json
import os
import subprocess
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from disco_b200 import gpu  # noqa: E402


def _edges(n, seed):
    rng = np.random.default_rng(seed)
    e = np.zeros(n, dtype=gpu.EDGE_DTYPE)
    e["src"] = rng.integers(0, 1 << 31, n); e["dst"] = rng.integers(0, 1 << 31, n)
    e["offset"] = rng.integers(1, 200, n); e["orient"] = rng.integers(0, 4, n)
    return e


def test_checksums_are_order_and_partition_independent():
    e = _edges(100_000, 1)
    whole = bench.edge_checksum(e)
    assert whole[0] == len(e)
    perm = np.random.default_rng(2).permutation(len(e))
    assert bench.edge_checksum(e[perm]) == whole
    cuts = [0, 1, 777, 50_000, 99_999, len(e)]
    parts = [bench.edge_checksum(e[perm][a:b]) for a, b in zip(cuts, cuts[1:])] + [bench.edge_checksum(e[:0])]
    assert bench.combine(parts) == whole                      # what the ranks of a multi-GPU run contribute
    # any change of one field of one edge changes the checksum; so does a dropped or doubled edge
    for field in ("src", "dst", "offset", "orient"):
        f = e.copy()
        f[field][12345] += 1
        assert bench.edge_checksum(f) != whole
    assert bench.edge_checksum(e[1:]) != whole
    assert bench.edge_checksum(np.concatenate([e, e[:1]])) != whole
    # swapping two columns of an edge is seen (the mix is not symmetric in its arguments)
    g = e.copy()
    g["src"], g["dst"] = e["dst"].copy(), e["src"].copy()
    assert bench.edge_checksum(g) != whole


def test_contained_row_checksum():
    rng = np.random.default_rng(3)
    c = np.zeros(5000, dtype=gpu.CROW_DTYPE)
    for k in ("contained", "container"):
        c[k] = rng.integers(0, 1 << 31, len(c))
    c["orient"] = rng.integers(0, 4, len(c)); c["start"] = rng.integers(0, 100, len(c))
    whole = bench.crow_checksum(c)
    assert bench.combine([bench.crow_checksum(c[:100]), bench.crow_checksum(c[100:])]) == whole
    d = c.copy(); d["container"][7] ^= 1
    assert bench.crow_checksum(d) != whole


def test_our_arm_refuses_to_run_without_a_gpu():
    """no CPU fallback: without a CUDA device bench.py's own arm exits with a message instead of timing anything"""
    import torch
    if torch.cuda.is_available():
        return
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and "no CUDA device" in (r.stdout + r.stderr)
    assert not any(l.startswith("{") and "value" in json.loads(l) for l in r.stdout.splitlines() if l.startswith("{"))
