"""buildG executable: CLI contract (CPU) and end-to-end file parity with the reference's outputs (GPU)."""
import os
import subprocess
import pytest
from helpers import GOLDEN, load_golden, HERE, check_partial_graphs

BUILDG = os.path.join(os.path.dirname(HERE), "disco_b200", "bin", "buildG")


def _run(args, **kw):
    return subprocess.run([BUILDG] + args, capture_output=True, text=True, **kw)


def test_usage_and_exit_codes(tmp_path):
    assert os.access(BUILDG, os.X_OK)
    r = _run([])                                   # main.cpp:93-102: no arguments -> usage, exit 0
    assert r.returncode == 0 and "Usage: buildG" in r.stderr
    r = _run(["-h"])
    assert r.returncode == 0 and "Usage: buildG" in r.stderr
    r = _run(["--bogus"])                          # main.cpp:133-148: unknown option -> usage, exit 1
    assert r.returncode == 1 and "Unknown option: --bogus" in r.stderr
    r = _run(["-se", "x.fa", "-f", str(tmp_path / "o"), "-p", str(tmp_path / "missing.cfg")])
    assert r.returncode == 1 and "Unable to open parameter file" in r.stderr   # main.cpp:156-159


def test_checkpoint_skips_finished_stage(tmp_path):
    cfg = tmp_path / "d.cfg"
    cfg.write_text("MinOverlap4BuildGraph = 30\n")
    (tmp_path / "o_CheckpointInfo.txt").write_text("CCR=Complete\nGC=Complete\n")
    r = _run(["-se", "nonexistent.fa", "-f", str(tmp_path / "o"), "-p", str(cfg)])
    assert r.returncode == 0 and "Graph already exists" in r.stdout      # main.cpp:48-52


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["fixture_contained_m30", "filter_mix_m40", "paired_800x2x250_m30", "dup_contained_3000_m50"])
def test_files_match_reference(tmp_path, name):
    g = load_golden([p for p in GOLDEN if name in p][0])
    fa = tmp_path / "reads.fa"
    fa.write_text("".join(f">{i + 1}\n{s}\n" for i, s in enumerate(g["records"])))
    cfg = tmp_path / "disco.cfg"
    cfg.write_text(f"# comment\nMinOverlap4BuildGraph = {g['min_overlap']}\nMinOverlap4SimplifyGraph = 40\n")
    prefix = str(tmp_path / "graph" / "out")
    os.makedirs(os.path.dirname(prefix))
    paired = "paired" in name
    r = _run(["-pe" if paired else "-se", str(fa), "-f", prefix, "-p", str(cfg), "-t", "3", "-m", "8"])
    assert r.returncode == 0, r.stdout + r.stderr
    assert "Graph construction complete." in r.stdout
    for t in range(3):   # every file runDisco.sh lists for -n 3 must exist
        for suffix in ("parGraph", "containedReads", "startRead"):
            assert os.path.exists(f"{prefix}_{t}_{suffix}.txt")
    # -t 3: three partial graphs with the reference's mark flags; their union is the reference's edge set
    assert check_partial_graphs(prefix, 3) == set(g["ref_edges"])
    if len(g["ref_edges"]) > 50:
        assert all(os.path.getsize(f"{prefix}_{t}_parGraph.txt") > 0 for t in range(3))
    rows = [l.rstrip("\n") for t in range(3) for l in open(f"{prefix}_{t}_containedReads.txt")]
    assert rows == g["ref_crows"]
    kind = "Paired-end" if paired else "Singleton"
    assert open(prefix + "_ReadIDMap.txt").read() == f"{fa}: {kind} file 1\nReadID Range: (1,{len(g['records'])})\n"
    assert open(prefix + "_CheckpointInfo.txt").read().split() == ["CCR=Complete", "GC=Complete"]
    r2 = _run(["-se", str(fa), "-f", prefix, "-p", str(cfg)])
    assert r2.returncode == 0 and "Graph already exists" in r2.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["paired_800x2x250_m30", "dup_contained_3000_m50"])
def test_several_gpus_in_one_process(tmp_path, name):
    """-g a,b,c = the key-sharded partitioning inside one process (disco_gpu_build_graph_multi).  With one GPU in the box
    the three contexts share device 0; the files must still equal the reference's."""
    g = load_golden([p for p in GOLDEN if name in p][0])
    fa = tmp_path / "reads.fa"
    fa.write_text("".join(f">{i + 1}\n{s}\n" for i, s in enumerate(g["records"])))
    cfg = tmp_path / "disco.cfg"
    cfg.write_text(f"MinOverlap4BuildGraph = {g['min_overlap']}\n")
    prefix = str(tmp_path / "graph" / "out")
    os.makedirs(os.path.dirname(prefix))
    import torch
    devs = "0,1,0" if torch.cuda.device_count() >= 2 else "0,0,0"
    r = _run(["-pe" if "paired" in name else "-se", str(fa), "-f", prefix, "-p", str(cfg), "-t", "2", "-g", devs])
    assert r.returncode == 0, r.stdout + r.stderr
    assert check_partial_graphs(prefix, 2) == set(g["ref_edges"])
    edges = sorted(l + ",2" for l in g["ref_edges"])
    rows = [l.rstrip("\n") for t in range(2) for l in open(f"{prefix}_{t}_containedReads.txt")]
    assert rows == g["ref_crows"]
    # -g all: every visible device
    prefix2 = str(tmp_path / "graph" / "all")
    r = _run(["-pe" if "paired" in name else "-se", str(fa), "-f", prefix2, "-p", str(cfg), "-g", "all"])
    assert r.returncode == 0, r.stdout + r.stderr
    assert sorted(l.rstrip("\n") for l in open(f"{prefix2}_0_parGraph.txt")) == edges
    r = _run(["-se", str(fa), "-f", str(tmp_path / "x"), "-p", str(cfg), "-g", "0,1,2,3,4,5,6,7,8"])
    assert r.returncode == 1 and "at most 8 GPUs" in r.stdout
