"""Multi-rank path on CPU: world_size 2 (and 3) over gloo.  Clusters are sharded
cluster i -> rank i % world, each rank writes its own batch files, rank 0 writes the
aggregate; everything must equal the single-process reference golden."""
import glob
import os
import socket
import subprocess
import sys

import pytest

from fandom_search_b200 import parallel
from tests.util import compare_records, read_csv

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _gpu_count():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.gpu
def test_sharded_analyze_on_two_gpus(golden_dir, tmp_path):
    """Same as below on real devices: 2 ranks, one B200 each, NCCL process group."""
    if _gpu_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    _run_world(golden_dir, tmp_path, 2, real=True)


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_analyze_equals_single_process_golden(golden_dir, tmp_path, world):
    _run_world(golden_dir, tmp_path, world, real=False)


def _run_world(golden_dir, tmp_path, world, real):
    os.symlink(os.path.join(golden_dir, "fanworks"), tmp_path / "fanworks")
    os.symlink(os.path.join(golden_dir, "script.txt"), tmp_path / "script.txt")
    port = _free_port()
    procs = []
    for rank in range(world):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                   MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        if real:
            env["FS_TEST_REAL_DEVICE"] = "1"
        else:
            env["CUDA_VISIBLE_DEVICES"] = ""
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "_dist_worker.py"),
                                       golden_dir, str(tmp_path)], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=600)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
    # each rank announced only its own clusters: the balanced table every rank computes from the
    # bytes of text per cluster (parallel.assign_clusters)
    import random
    listing = open(os.path.join(golden_dir, "listing.txt")).read().split()
    files = [os.path.join(golden_dir, "fanworks", f) for f in listing]
    random.seed(4815162342)
    random.shuffle(files)
    sizes = [sum(os.path.getsize(f) for f in files[i:i + 16]) for i in range(0, len(files), 16)]
    owner = parallel.assign_clusters(sizes, world)
    assert len(sizes) == 3 and sorted(set(owner)) == list(range(min(world, 3)))
    for rank, o in enumerate(outs):
        seen = [int(l.split()[2]) for l in o.splitlines() if l.startswith("Processing cluster")]
        assert seen == [i for i in range(3) if owner[i] == rank]
    aggs = glob.glob(str(tmp_path / "match-6gram-2*.csv"))
    assert len(aggs) == 1                                   # only rank 0 writes the aggregate
    got = read_csv(aggs[0])
    want = read_csv(os.path.join(golden_dir, "golden_exhaustive.csv"))
    compare_records(got, want, tol=1e-12, basename=False)
    assert [(r[0], r[1]) for r in got] == [(r[0], r[1]) for r in want]
    for i in range(3):
        b = read_csv(str(tmp_path / ("match-6gram-batch-%d.csv" % i)), header=False)
        wb = read_csv(os.path.join(golden_dir, "golden_exhaustive.batch%d.csv" % i), header=False)
        assert [(r[0], r[1]) for r in b] == [(r[0], r[1]) for r in wb]


def test_cluster_owner_is_round_robin():
    assert [parallel.cluster_owner(i, 4) for i in range(9)] == [0, 1, 2, 3, 0, 1, 2, 3, 0]
    assert parallel.assign_clusters([5] * 9, 4, policy='roundrobin') == [0, 1, 2, 3, 0, 1, 2, 3, 0]


def test_balanced_assignment_leaves_no_tail():
    # equal clusters: the same loads as round robin; a short last cluster and one heavy cluster: LPT
    assert sorted(parallel.assign_clusters([7] * 8, 4)) == [0, 0, 1, 1, 2, 2, 3, 3]
    sizes = [10, 10, 10, 10, 10, 10, 10, 10, 3]          # 9 clusters on 4 ranks
    owner = parallel.assign_clusters(sizes, 4)
    load = [sum(s for s, o in zip(sizes, owner) if o == r) for r in range(4)]
    assert max(load) == 23 and max(load) - min(load) <= 3 and len(owner) == 9
    rr = [sum(s for i, s in enumerate(sizes) if i % 4 == r) for r in range(4)]
    assert max(rr) == 23                                  # here round robin happens to be as good ...
    sizes = [30, 5, 5, 5, 30, 5, 5, 5]                    # ... here it puts both heavy clusters on rank 0
    owner = parallel.assign_clusters(sizes, 4)
    load = [sum(s for s, o in zip(sizes, owner) if o == r) for r in range(4)]
    assert max(load) == 30 and max(sum(s for i, s in enumerate(sizes) if i % 4 == r) for r in range(4)) == 60
    assert parallel.assign_clusters(sizes, 4) == owner and parallel.assign_clusters(sizes, 1) == [0] * 8


def test_oov_hash_is_rank_independent_under_torchrun(monkeypatch):
    from fandom_search_b200 import search
    from fandom_search_b200.lexicon import py_hash_seed0
    monkeypatch.delenv("FANDOM_SEARCH_OOV_HASH", raising=False)
    monkeypatch.delenv("PYTHONHASHSEED", raising=False)
    monkeypatch.setenv("WORLD_SIZE", "1")
    assert search._default_oov_hash() is None                 # one interpreter: the builtin hash, as the reference
    monkeypatch.setenv("WORLD_SIZE", "8")
    assert search._default_oov_hash() is py_hash_seed0        # separate interpreters: one hash for all ranks
    monkeypatch.setenv("PYTHONHASHSEED", "123")
    assert search._default_oov_hash() is None                 # the environment pinned the builtin hash itself
    monkeypatch.setenv("FANDOM_SEARCH_OOV_HASH", "seed0")
    assert search._default_oov_hash() is py_hash_seed0


def test_a_failing_rank_stops_the_others(golden_dir, tmp_path):
    """One rank cannot read its cluster: every rank must end with an error promptly (the failure is
    all_reduced before the barrier) instead of hanging until the collective times out."""
    os.symlink(os.path.join(golden_dir, "fanworks"), tmp_path / "fanworks")
    os.symlink(os.path.join(golden_dir, "script.txt"), tmp_path / "script.txt")
    port = _free_port()
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), CUDA_VISIBLE_DEVICES="", FS_TEST_FAIL_RANK="1")
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "_dist_worker.py"),
                                       golden_dir, str(tmp_path)], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=120)[0] for p in procs]
    assert procs[0].returncode != 0 and procs[1].returncode != 0
    assert "simulated unreadable fanwork" in outs[1]
    assert "another rank failed" in outs[0]
    assert not glob.glob(str(tmp_path / "match-6gram-2*.csv"))       # no aggregate from a failed run
